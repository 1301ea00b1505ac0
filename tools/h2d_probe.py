"""H2D bandwidth from pinned host memory on this box: one 36 MB copy, the bench's 12 + 24 MB split, and two streams."""
import torch
dev = torch.device("cuda", 0)
def bw(sizes, streams=1, reps=50):
    hs = [torch.empty(s, dtype=torch.uint8).pin_memory() for s in sizes]
    ds = [torch.empty(s, dtype=torch.uint8, device=dev) for s in sizes]
    st = [torch.cuda.Stream(dev) for _ in range(streams)]
    def go():
        for i, (h, d) in enumerate(zip(hs, ds)):
            with torch.cuda.stream(st[i % streams]):
                d.copy_(h, non_blocking=True)
    for _ in range(5): go()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for s in st: s.wait_stream(torch.cuda.current_stream(dev))
    for _ in range(reps): go()
    for s in st: torch.cuda.current_stream(dev).wait_stream(s)
    b.record(); torch.cuda.synchronize()
    return sum(sizes) * reps / (a.elapsed_time(b) / 1e3) / 1e9
M = 1 << 20
print(f"one 36.2 MB copy            : {bw([36246528]):.1f} GB/s")
print(f"12.1 + 24.2 MB (bench split): {bw([12082176, 24164352]):.1f} GB/s")
print(f"same on two streams         : {bw([12082176, 24164352], streams=2):.1f} GB/s")
print(f"one 256 MB copy             : {bw([256 * M], reps=10):.1f} GB/s")
