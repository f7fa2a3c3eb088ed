"""Wide bf16 transform with the training epilogues (dropout / keep-mask) vs the plain one."""
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
a1 = torch.randn(M, 1024, device=dev).to(bf); a2 = torch.randn(M, 1024, device=dev).to(bf)
w1 = (torch.randn(1024, 1024, device=dev) / 32).to(bf); w2 = (torch.randn(1024, 1024, device=dev) / 32).to(bf)
b = torch.randn(1024, device=dev); act = torch.randn(M, 1024, device=dev).to(bf)
print(f"plain   : {timeit(lambda: ops.node_transform(a1, w1, a2, w2, b, relu=True))*1e3:.1f} us")
print(f"dropout : {timeit(lambda: ops.node_transform(a1, w1, a2, w2, b, relu=True, dropout=(0.1, 7)))*1e3:.1f} us")
print(f"mask    : {timeit(lambda: ops.node_transform(a1, w1, a2, w2, None, relu=False, mask=(act, 1.1)))*1e3:.1f} us")
