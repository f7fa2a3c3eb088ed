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
opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
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
