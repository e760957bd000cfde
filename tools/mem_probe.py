"""Measures plain device-memory rates on this GPU (write-only, read-only, copy) to put the aggregate kernel's
bound in context: fill = torch memset-like kernel, read = sum reduction, copy = clone."""
import torch
dev = "cuda"
n = 1 << 28                                   # 1 GiB of fp32
x = torch.empty(n, device=dev)
y = torch.empty(n, device=dev)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e-3
gb = n * 4 / 1e9
t = timeit(lambda: x.fill_(1.0)); print(f"fill  (write only): {gb / t:8.0f} GB/s")
t = timeit(lambda: x.sum());      print(f"sum   (read only) : {gb / t:8.0f} GB/s")
t = timeit(lambda: y.copy_(x));   print(f"copy  (read+write): {2 * gb / t:8.0f} GB/s")
t = timeit(lambda: torch.add(x, 1.0, out=y)); print(f"add   (read+write): {2 * gb / t:8.0f} GB/s")
for mb in (26, 105, 209):
    m = mb * (1 << 20) // 4
    t = timeit(lambda: x[:m].fill_(1.0), 20); print(f"fill {mb:4d} MB: {m * 4 / 1e9 / t:8.0f} GB/s ({t * 1e6:.1f} us)")
    t = timeit(lambda: y[:m].copy_(x[:m]), 20); print(f"copy {mb:4d} MB: {2 * m * 4 / 1e9 / t:8.0f} GB/s ({t * 1e6:.1f} us)")
