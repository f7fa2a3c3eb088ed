"""Write-only / read-only / copy bandwidth of this GPU (context for write-dominated kernels)."""
import torch
dev = torch.device("cuda:0")
def timeit(f, n=20):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
for mb in (205, 1024, 4096):
    n = mb * (1 << 20) // 2
    x = torch.empty(n, dtype=torch.bfloat16, device=dev); y = torch.empty_like(x)
    t = timeit(lambda: x.zero_());      print(f"{mb} MB fill : {t*1e3:8.1f} us  {mb*1.048576/t:8.1f} GB/s")
    t = timeit(lambda: y.copy_(x));     print(f"{mb} MB copy : {t*1e3:8.1f} us  {2*mb*1.048576/t:8.1f} GB/s (read+write)")
    t = timeit(lambda: x.view(torch.int16).max()); print(f"{mb} MB read (torch max)  : {t*1e3:8.1f} us  {mb*1.048576/t:8.1f} GB/s")
