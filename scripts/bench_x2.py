"""fp32 node transform on the tensor cores (lpgnn_node_transform_x2): time and accuracy by chunk length on the C2 hidden
layer shapes, next to the CUDA-core fp32 kernel, the 16-bit transform and the split pass that feeds it."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200  # noqa: F401
from lpgnn_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(f, n=10):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
torch.manual_seed(0)
N = K = 1024
w1 = torch.randn(N, K, device=dev) / 32; w2 = torch.randn(N, K, device=dev) / 32
b = torch.randn(N, device=dev)
hw = torch.randn(3, N, device=dev) / 32; hb = torch.randn(3, device=dev)
pw1, pw2, cs = ops.split_x2(w1, w2)
for M in (100_000, 50_000, 4096):
    a1 = torch.randn(M, K, device=dev).relu(); a2 = torch.randn(M, K, device=dev).relu()
    x = torch.randn(M, 8, device=dev)
    pa1, pa2, rs = ops.split_x2(a1, a2)
    t_split = timeit(lambda: ops.split_x2(a1, a2))
    e = None
    if M <= 4096:
        e = a1.double() @ w1.double().T + a2.double() @ w2.double().T + b.double()
    fl = 4.0 * M * N * K
    print(f"M={M}: split_x2 {t_split*1e3:.1f} us ({M*2*K*8/t_split/1e6:.0f} GB/s moved)", flush=True)
    for ck in (1, 2, 4, 8, 16):
        ops.set_x2_chunk(ck)
        t = timeit(lambda: ops.node_transform_x2(pa1, pw1, pa2, pw2, rs, cs, b, relu=False))
        th = timeit(lambda: ops.node_transform_x2(pa1, pw1, pa2, pw2, rs, cs, b, relu=True, head=(hw, hb, x), want_out=False))
        msg = f"  chunk {ck:2d}: {t*1e3:7.1f} us = {3*fl/t/1e9:6.0f} half-TF/s ({fl/t/1e9:5.0f} fp32-equivalent) | head-fused, no store {th*1e3:7.1f} us"
        if e is not None:
            y = ops.node_transform_x2(pa1, pw1, pa2, pw2, rs, cs, b)
            msg += f" | rel fro err {float((y.double()-e).norm()/e.norm()):.2e} max/rowmax {float(((y.double()-e).abs()/e.abs().amax(1,keepdim=True)).max()):.2e}"
        print(msg, flush=True)
    ops.set_x2_chunk(4)
    if M <= 50_000:
        t = timeit(lambda: ops.node_transform(a1, w1, a2, w2, b), 3)
        msg = f"  CUDA-core fp32 kernel: {t*1e3:.1f} us ({fl/t/1e9:.0f} TF/s)"
        if e is not None:
            y = ops.node_transform(a1, w1, a2, w2, b)
            msg += f" | rel fro err {float((y.double()-e).norm()/e.norm()):.2e} max/rowmax {float(((y.double()-e).abs()/e.abs().amax(1,keepdim=True)).max()):.2e}"
        print(msg, flush=True)
    h = torch.float16
    a1h, a2h, w1h, w2h = a1.to(h), a2.to(h), w1.to(h), w2.to(h)
    t = timeit(lambda: ops.node_transform(a1h, w1h, a2h, w2h, b))
    print(f"  16-bit (half) transform, one pass: {t*1e3:.1f} us ({fl/t/1e9:.0f} TF/s)", flush=True)
    del a1, a2, pa1, pa2, a1h, a2h
