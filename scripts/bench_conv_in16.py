"""Input layer in the 16-bit modes (lpgnn_conv_in_16 / lpgnn_conv_in_16_pair) on a C2-shaped LP: both directions, time and fraction of the write roofline."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200  # noqa: F401
from lpgnn_b200 import arch, ops, synth
from lpgnn_b200.graph import BipartiteCSR
dev = torch.device("cuda:0")
cfg = synth.CONFIGS["C2"]
lp = synth.processed_lp(cfg["m"], cfg["n"], cfg["nnz"], seed=cfg["seed"])
g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n, dev, is_sorted=True)
csr, csc = g.views()
xs, xt = torch.from_numpy(lp.c_feas).to(dev), torch.from_numpy(lp.v_feas).to(dev)
torch.manual_seed(0)
model = arch.GCN_FC(8, 8, hids=1024, depth=3).to(dev)
c1 = model.conv1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(f, n=20):
    for _ in range(3): f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2]
w = lambda p: p.detach()
for dt in (torch.float16, torch.bfloat16):
    ft = lambda: ops.conv_in_16(csc, xs, xt, w(c1.left2right.lin_rel.weight), w(c1.left2right.lin_rel.bias), w(c1.left2right.lin_root.weight), dt)
    fs = lambda: ops.conv_in_16(csr, xt, xs, w(c1.right2left.lin_rel.weight), w(c1.right2left.lin_rel.bias), w(c1.right2left.lin_root.weight), dt)
    l2r = (w(c1.left2right.lin_rel.weight), w(c1.left2right.lin_rel.bias), w(c1.left2right.lin_root.weight))
    r2l = (w(c1.right2left.lin_rel.weight), w(c1.right2left.lin_rel.bias), w(c1.right2left.lin_root.weight))
    fp = lambda: ops.conv_in_16_pair(csr, csc, xs, xt, l2r, r2l, dt)
    tt, ts_, tp = timeit(ft), timeit(fs), timeit(fp)
    by = (lp.m + lp.n) * 1024 * 2
    print(f"{dt}: both directions in ONE launch {tp*1e3:.1f} us = {by/tp/1e6:.0f} GB/s written ({by/tp/1e6/6538.9:.2f})", flush=True)
    print(f"{dt}: vars side {tt*1e3:.1f} us, cons side {ts_*1e3:.1f} us, pair {1e3*(tt+ts_):.1f} us = {by/(tt+ts_)/1e6:.0f} GB/s written "
          f"({by/(tt+ts_)/1e6/6538.9:.2f} of the measured copy bandwidth)", flush=True)
