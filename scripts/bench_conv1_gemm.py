"""Times the one-K-block input transform and the hidden transform (diagnostic)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200
from lpgnn_b200 import ops
dev = torch.device("cuda:0")
bf = torch.bfloat16
def timeit(f, n=30):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]
for M in (100_000, 50_000):
    z = torch.randn(M, 64, device=dev).to(bf); w = torch.randn(1024, 64, device=dev).to(bf); b = torch.randn(1024, device=dev)
    print(f"conv1 GEMM M={M}: {timeit(lambda: ops.node_transform(z, w, bias=b, relu=True))*1e3:.1f} us")
    a1 = torch.randn(M, 1024, device=dev).to(bf); a2 = torch.randn(M, 1024, device=dev).to(bf)
    w1 = torch.randn(1024, 1024, device=dev).to(bf); w2 = torch.randn(1024, 1024, device=dev).to(bf)
    print(f"hidden GEMM M={M}: {timeit(lambda: ops.node_transform(a1, w1, a2, w2, b, relu=True))*1e3:.1f} us")
    hw = torch.randn(3, 1024, device=dev); hb = torch.randn(3, device=dev); x = torch.randn(M, 8, device=dev)
    print(f"hidden GEMM + head M={M}: {timeit(lambda: ops.node_transform_head(a1, w1, a2, w2, b, hw, hb, x))*1e3:.1f} us")
