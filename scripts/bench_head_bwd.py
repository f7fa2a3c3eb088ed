"""head_mask_bwd with the fused column sums (bias gradient of the layer under the head) vs head_mask_bwd + colsum."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200
from lpgnn_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(f, n=20):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
for dtype in (torch.bfloat16, torch.float32):
    for rows in (100_000, 50_000):
        H = 1024
        g = torch.Generator(device="cuda").manual_seed(1)
        h = torch.randn(rows, H, device=dev, generator=g).relu().to(dtype)
        w = torch.randn(3, H, device=dev, generator=g) / 32
        raw = torch.randn(rows, 3, device=dev, generator=g); dl = torch.randn(rows, 3, device=dev, generator=g)
        dH = ops.head_mask_bwd(dl, raw, h, w, 1.1)[0]
        t0 = timeit(lambda: ops.head_mask_bwd(dl, raw, h, w, 1.1, want_bf16=dtype == torch.bfloat16))
        t1 = timeit(lambda: ops.colsum(dH))
        t2 = timeit(lambda: ops.head_mask_bwd(dl, raw, h, w, 1.1, want_bf16=dtype == torch.bfloat16, want_colsum=True))
        byts = rows * H * 2 * h.element_size()
        print(f"{dtype} rows={rows}: head_mask_bwd {t0*1e3:.1f} us ({byts/t0/1e6:.0f} GB/s) + colsum {t1*1e3:.1f} us = {(t0+t1)*1e3:.1f} us; fused {t2*1e3:.1f} us ({byts/t2/1e6:.0f} GB/s)")
