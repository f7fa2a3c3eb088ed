"""Graph build (sorted / unsorted COO -> CSR + CSC) on a C2-shaped LP and on a small one: time per build, L2 flushed."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200  # noqa: F401
from lpgnn_b200 import synth
from lpgnn_b200.graph import BipartiteCSR
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
for name, (m, n, z) in (("C2", (50_000, 100_000, 500_000)), ("small", (2_000, 4_000, 20_000)), ("C4/4", (250_000, 500_000, 2_500_000))):
    lp = synth.processed_lp(m, n, z, seed=3)
    t = lambda a, dt: torch.from_numpy(a.astype(dt)).to(dev)
    row, col, val = t(lp.row, np.int32), t(lp.col, np.int32), t(lp.a_data, np.float32)
    perm = torch.randperm(row.numel(), device=dev)
    ts = timeit(lambda: BipartiteCSR.from_coo(row, col, val, lp.m, lp.n, is_sorted=True))
    tu = timeit(lambda: BipartiteCSR.from_coo(row[perm], col[perm], val[perm], lp.m, lp.n, is_sorted=False))
    print(f"{name:6s} nnz {lp.nnz:8d}: sorted {ts*1e3:7.1f} us, unsorted {tu*1e3:7.1f} us", flush=True)
