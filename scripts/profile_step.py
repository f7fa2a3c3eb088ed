"""Exactly one C2 bf16 prediction step (lpgnn_predict_basis: graph build + GCN_FC forward + basis selection) inside a
cudaProfilerStart/Stop window, after warm-up -- the target of the ncu captures under profiles/ (run ncu with
--profile-from-start off).  `train` as first argument profiles one training step (forward + loss + backward) instead; the optional second argument is
the precision (bf16 | fp16)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200  # noqa: F401
from lpgnn_b200 import arch, synth
from lpgnn_b200.pipeline import pack_lp, unpack_device

mode = sys.argv[1] if len(sys.argv) > 1 else "predict"
precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"       # bf16 | fp16 | fp32 (fp16: prediction only)
dev = torch.device("cuda:0")
cfg = synth.CONFIGS["C2"]
lp = synth.processed_lp(cfg["m"], cfg["n"], cfg["nnz"], seed=cfg["seed"], structure="staircase")
torch.manual_seed(0)
model = arch.GCN_FC(8, 8, hids=cfg["hids"], depth=cfg["depth"]).to(dev)
model.set_precision(precision)
host_lp = pack_lp(lp.row, lp.col, lp.a_data, lp.c_feas, lp.v_feas, is_sorted=True)
d_row, d_col, d_val, d_xs, d_xt = (t.clone() for t in unpack_device(host_lp.pack.to(dev), host_lp))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

if mode == "predict":
    model.eval()
    step = lambda: model.predict_basis_coo(d_row, d_col, d_val, lp.m, lp.n, d_xs, d_xt, is_sorted=True)
else:
    from lpgnn_b200.data import Data
    from lpgnn_b200.graph import BipartiteCSR
    from lpgnn_b200 import losses
    model.train()
    g = BipartiteCSR.from_coo(d_row, d_col, d_val, lp.m, lp.n, is_sorted=True)
    batch = Data(x_s=d_xs, x_t=d_xt, edge_index=g)
    y_s, y_t = torch.from_numpy(lp.y_s).to(dev), torch.from_numpy(lp.y_t).to(dev)

    def step():
        lc, lv = model(batch)
        loss = losses.balanced(lc, lv, y_s, y_t)
        model.zero_grad(set_to_none=True)
        loss.backward()
        return loss

for _ in range(5):
    step()
torch.cuda.synchronize()
flush.zero_()
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", mode)
