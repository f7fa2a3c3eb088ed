"""Input layer (conv1) on a C2-shaped LP: SIMT fused path (lpgnn_conv_in_fused) vs gather_cat + one-K-block tcgen05 transform."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200
from lpgnn_b200 import ops, synth
from lpgnn_b200.graph import BipartiteCSR
dev = torch.device("cuda:0")
lp = synth.config_lp("C2")
g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype("float32"), lp.m, lp.n, dev, is_sorted=True)
csr, csc = g.views()
xs, xt = torch.from_numpy(lp.c_feas).to(dev), torch.from_numpy(lp.v_feas).to(dev)
H = 1024
torch.manual_seed(0)
w_rel, w_root, b = torch.randn(H, 8, device=dev), torch.randn(H, 8, device=dev), torch.randn(H, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(f, n=20):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e))
    ts.sort(); return ts[len(ts) // 2]
wcat = torch.zeros(H, 64, device=dev); wcat[:, :8] = w_rel; wcat[:, 8:16] = w_root
wcat_b = wcat.to(torch.bfloat16)
for name, view, src, dst in (("vars side", csc, xs, xt), ("cons side", csr, xt, xs)):
    t1 = timeit(lambda: ops.conv_in_fused(view, src, dst, w_rel, b, w_root, torch.bfloat16))
    t2 = timeit(lambda: ops.gather_cat(view, src, dst, want_f32=False, want_bf16=True))
    _, zb = ops.gather_cat(view, src, dst, want_f32=False, want_bf16=True)
    t3 = timeit(lambda: ops.node_transform(zb, wcat_b, bias=b, relu=True))
    rows = view[3]
    print(f"{name} ({rows} rows): SIMT fused (gather + FFMA2 transform) {t1*1e3:.1f} us | gather_cat {t2*1e3:.1f} us + tcgen05 one-K-block {t3*1e3:.1f} us"
          f" | output {rows*H*2/1e6:.0f} MB -> fill-speed floor ~{rows*H*2/3.4e6:.0f} us")
