"""Host-side enqueue time vs GPU time of each stage of one basis-prediction step (diagnostic)."""
import os, sys, time, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200
from lpgnn_b200 import arch, ops, synth, _lib
from lpgnn_b200.graph import BipartiteCSR

dev = torch.device("cuda:0")
lp = synth.config_lp("C2")
m, n = lp.m, lp.n
row = torch.from_numpy(lp.row.astype(np.int32)).to(dev); col = torch.from_numpy(lp.col.astype(np.int32)).to(dev)
val = torch.from_numpy(lp.a_data.astype(np.float32)).to(dev)
xs = torch.from_numpy(lp.c_feas).to(dev); xt = torch.from_numpy(lp.v_feas).to(dev)
torch.manual_seed(0)
model = arch.GCN_FC(8, 8, hids=1024, depth=3).to(dev).eval().set_precision("bf16")

def stage(name, fn, reps=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    t_cpu = (time.perf_counter() - t0) / reps
    torch.cuda.synchronize()
    t_all = (time.perf_counter() - t0) / reps
    worst = 0
    for _ in range(reps):
        t1 = time.perf_counter(); fn(); worst = max(worst, time.perf_counter() - t1)
    torch.cuda.synchronize()
    print(f"{name:28s} enqueue {t_cpu*1e3:7.3f} ms   enqueue+drain {t_all*1e3:7.3f} ms   worst enqueue {worst*1e3:7.3f} ms")

g = BipartiteCSR.from_coo(row, col, val, m, n, is_sorted=True)
batch = types.SimpleNamespace(x_s=xs, x_t=xt, edge_index=g)
with torch.no_grad():
    stage("graph build (sorted)", lambda: BipartiteCSR.from_coo(row, col, val, m, n, is_sorted=True))
    stage("graph build (unsorted)", lambda: BipartiteCSR.from_coo(row, col, val, m, n, is_sorted=False))
    stage("forward", lambda: model(batch))
    lc, lv = model(batch)
    stage("basis_select", lambda: ops.basis_select(lc, lv, int64=False))
    def full():
        gg = BipartiteCSR.from_coo(row, col, val, m, n, is_sorted=True)
        return model.predict_basis(types.SimpleNamespace(x_s=xs, x_t=xt, edge_index=gg), int64=False)
    stage("full step", full)
    lib = _lib.load()
    i32 = dict(dtype=torch.int32, device=dev)
    z = row.shape[0]
    outs = [torch.empty(m + 1, **i32), torch.empty(z, **i32), torch.empty(z, dtype=torch.float32, device=dev),
            torch.empty(n + 1, **i32), torch.empty(z, **i32), torch.empty(z, dtype=torch.float32, device=dev), torch.empty(z, **i32)]
    status = torch.zeros(1, **i32)
    nb = lib.lpgnn_graph_build_workspace_bytes(z, m, n)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    sp = torch.cuda.current_stream().cuda_stream
    def raw():
        lib.lpgnn_graph_build(row.data_ptr(), col.data_ptr(), 0, val.data_ptr(), z, m, n, 1, *[t.data_ptr() for t in outs],
                              status.data_ptr(), ws.data_ptr(), nb, sp)
    stage("raw lpgnn_graph_build call", raw)
