import os, sys, time, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200
from lpgnn_b200 import arch, ops, synth, _lib
from lpgnn_b200.graph import BipartiteCSR
dev = torch.device("cuda:0")
lp = synth.config_lp("C2"); m, n = lp.m, lp.n
row = torch.from_numpy(lp.row.astype(np.int32)).to(dev); col = torch.from_numpy(lp.col.astype(np.int32)).to(dev)
val = torch.from_numpy(lp.a_data.astype(np.float32)).to(dev)
xs = torch.from_numpy(lp.c_feas).to(dev); xt = torch.from_numpy(lp.v_feas).to(dev)
torch.manual_seed(0)
model = arch.GCN_FC(8, 8, hids=1024, depth=3).to(dev).eval().set_precision("bf16")
def st():
    s = torch.cuda.memory_stats()
    return s["num_device_alloc"], s["num_device_free"], s["num_alloc_retries"], s["reserved_bytes.all.current"] >> 20
def full():
    gg = BipartiteCSR.from_coo(row, col, val, m, n, is_sorted=True)
    return model.predict_basis(types.SimpleNamespace(x_s=xs, x_t=xt, edge_index=gg), int64=False)
def build():
    return BipartiteCSR.from_coo(row, col, val, m, n, is_sorted=True)
with torch.no_grad():
    for name, fn in (("build", build), ("full", full), ("build", build), ("full", full)):
        for sync in (False, True):
            for _ in range(3): fn()
            torch.cuda.synchronize(); s0 = st()
            ts = []
            t0 = time.perf_counter()
            for i in range(40):
                t1 = time.perf_counter(); fn()
                if sync: torch.cuda.synchronize()
                ts.append(time.perf_counter() - t1)
            torch.cuda.synchronize()
            tot = (time.perf_counter() - t0) / 40
            ts = np.array(ts) * 1e3
            print(f"{name:6s} sync={sync!s:5s} avg {tot*1e3:7.3f} ms  per-call median {np.median(ts):6.3f} max {ts.max():7.3f} n>1ms {int((ts>1).sum())}  alloc stats before {s0} after {st()}")
