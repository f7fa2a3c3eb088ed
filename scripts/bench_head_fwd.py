"""head_mask forward (basis-status head + knowledge mask) on C2-sized activations: time and fraction of the HBM roofline."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200
from lpgnn_b200 import _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
lib = _lib.load()
def timeit(f, n=20):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
for dtype in (torch.bfloat16, torch.float16, torch.float32):
    for rows in (100_000, 50_000):
        H = 1024
        g = torch.Generator(device="cuda").manual_seed(1)
        h = torch.randn(rows, H, device=dev, generator=g).relu().to(dtype)
        w = torch.randn(3, H, device=dev, generator=g) / 32
        b = torch.randn(3, device=dev, generator=g)
        feas = torch.randint(-1, 2, (rows, 8), device=dev, generator=g).float()
        logits = torch.empty(rows, 3, device=dev); raw = torch.empty(rows, 3, device=dev)
        st = _lib.stream_ptr()
        f = lambda: lib.lpgnn_head_mask(h.data_ptr(), _lib.dtype_code(dtype), rows, H, w.data_ptr(), b.data_ptr(), feas.data_ptr(), 8,
                                        logits.data_ptr(), raw.data_ptr(), st)
        t = timeit(f)
        byts = rows * (H * h.element_size() + 8 * 4 + 24)
        print(f"{dtype} rows={rows}: head_mask {t*1e3:.1f} us = {byts/t/1e6:.0f} GB/s ({byts/t/1e6/6538.9:.3f} of HBM)")
