"""A/B of the wide bf16 node transform: 2-CTA clusters with multicast W tiles vs one CTA per tile (lpgnn_set_gemm_cluster)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200
from lpgnn_b200 import _lib, ops
dev = torch.device("cuda:0")
bf = torch.bfloat16
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(f, n=20):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
torch.manual_seed(0)
for M in (100_000, 50_000, 100_001, 2_000_000):
    a1 = torch.randn(M, 1024, device=dev).to(bf); a2 = torch.randn(M, 1024, device=dev).to(bf)
    w1 = (torch.randn(1024, 1024, device=dev) / 32).to(bf); w2 = (torch.randn(1024, 1024, device=dev) / 32).to(bf)
    b = torch.randn(1024, device=dev)
    hw = torch.randn(3, 1024, device=dev); hb = torch.randn(3, device=dev); x = torch.randn(M, 8, device=dev)
    res = {}
    for mode in (0, 1):
        lib.lpgnn_set_gemm_cluster(mode)
        out = ops.node_transform(a1, w1, a2, w2, b, relu=True)
        lg = ops.node_transform_head(a1, w1, a2, w2, b, hw, hb, x)
        lg = lg[0] if isinstance(lg, tuple) else lg
        t = timeit(lambda: ops.node_transform(a1, w1, a2, w2, b, relu=True), 10 if M > 1_000_000 else 20)
        th = timeit(lambda: ops.node_transform_head(a1, w1, a2, w2, b, hw, hb, x), 10 if M > 1_000_000 else 20)
        res[mode] = (out, lg, t, th)
    same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    fl = 4.0 * M * 1024 * 1024
    print(f"M={M}: one CTA/tile {res[0][2]*1e3:.1f} us ({fl/res[0][2]/1e9:.0f} TF/s), head-fused {res[0][3]*1e3:.1f} us | "
          f"2-CTA clusters {res[1][2]*1e3:.1f} us ({fl/res[1][2]/1e9:.0f} TF/s), head-fused {res[1][3]*1e3:.1f} us | identical={same}", flush=True)
    del a1, a2, out, lg, res
lib.lpgnn_set_gemm_cluster(1)

# yardstick: the library GEMM (cuBLAS through torch.matmul) on the same shape and under the same timing protocol
for M in (100_000, 50_000):
    a = torch.randn(M, 2048, device=dev).to(bf)
    w = (torch.randn(1024, 2048, device=dev) / 32).to(bf)
    t = timeit(lambda: torch.matmul(a, w.t()))
    print(f"M={M}: cuBLAS (torch.matmul, K=2048 concatenated, no bias/ReLU epilogue) {t*1e3:.1f} us ({4.0*M*1024*1024/t/1e9:.0f} TF/s)")
a = torch.randn(8192, 8192, device=dev).to(bf); w = torch.randn(8192, 8192, device=dev).to(bf)
t = timeit(lambda: torch.matmul(a, w))
print(f"8192^3: cuBLAS {t*1e3:.1f} us ({2.0*8192**3/t/1e9:.0f} TF/s)")
