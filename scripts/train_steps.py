"""A few C2-shaped training steps (diagnostic driver for ncu launch lists)."""
import os, sys, types
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200
from lpgnn_b200 import arch, synth
from lpgnn_b200.graph import BipartiteCSR
from lpgnn_b200.losses import balanced
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "C2"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = synth.CONFIGS[name]
lp = synth.config_lp(name)
torch.manual_seed(0)
model = arch.GCN_FC(8, 8, hids=cfg["hids"], depth=cfg["depth"]).to(dev).train().set_precision(prec)
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4, fused=True)
g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n, dev, is_sorted=True)
batch = types.SimpleNamespace(x_s=torch.from_numpy(lp.c_feas).to(dev), x_t=torch.from_numpy(lp.v_feas).to(dev), edge_index=g)
y_s, y_t = torch.from_numpy(lp.y_s).to(dev), torch.from_numpy(lp.y_t).to(dev)
for i in range(steps):
    lc, lv = model(batch)
    loss = balanced(lc, lv, y_s, y_t)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
torch.cuda.synchronize()
print("loss", float(loss))
import time
def step():
    lc, lv = model(batch)
    loss = balanced(lc, lv, y_s, y_t)
    opt.zero_grad(set_to_none=True)
    loss.backward()
    opt.step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(20): step()
t_cpu = (time.perf_counter() - t0) / 20
torch.cuda.synchronize(); t_all = (time.perf_counter() - t0) / 20
print(f"train step: enqueue {t_cpu*1e3:.3f} ms, enqueue+drain {t_all*1e3:.3f} ms")
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=30, max_name_column_width=70))
