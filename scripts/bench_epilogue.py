"""Where does the one-K-block transform's time go?  Same GEMM with and without the activation store (head-only mode)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200
from lpgnn_b200 import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(f, n=20):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
M = 100_000
z = torch.randn(M, 64, device=dev).to(bf); w = torch.randn(1024, 64, device=dev).to(bf); b = torch.randn(1024, device=dev)
hw = torch.randn(3, 1024, device=dev); hb = torch.randn(3, device=dev); x = torch.randn(M, 8, device=dev)
print(f"K=64 with store      : {timeit(lambda: ops.node_transform(z, w, bias=b, relu=True))*1e3:.1f} us")
z2 = torch.randn(M, 64, device=dev).to(bf); w2 = torch.randn(1024, 64, device=dev).to(bf)
print(f"K=64+64 with store   : {timeit(lambda: ops.node_transform(z, w, z2, w2, b, relu=True))*1e3:.1f} us")
print(f"K=64+64 head-only (no activation store): {timeit(lambda: ops.node_transform_head(z, w, z2, w2, b, hw, hb, x, want_out=False))*1e3:.1f} us")
print(f"K=64+64 head + store : {timeit(lambda: ops.node_transform_head(z, w, z2, w2, b, hw, hb, x, want_out=True))*1e3:.1f} us")
for K in (128, 256, 512):
    a = torch.randn(M, K, device=dev).to(bf); wk = torch.randn(1024, K, device=dev).to(bf)
    print(f"K={K} with store     : {timeit(lambda: ops.node_transform(a, wk, bias=b, relu=True))*1e3:.1f} us")
