import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
import lpgnn_b200
from lpgnn_b200 import arch, synth
from lpgnn_b200.graph import BipartiteCSR
from lpgnn_b200.losses import balanced
from lpgnn_b200.sampling import NeighborSubgraphLoader, ResidentLP
dev = torch.device("cuda:0")
cfg = synth.CONFIGS["C3"]
lp = synth.processed_lp(cfg["m"], cfg["n"], cfg["nnz"], seed=cfg["seed"])
torch.manual_seed(0)
model = arch.GCN_FC(8, 8, hids=cfg["hids"], depth=cfg["depth"]).to(dev).train().set_precision("bf16")
params = list(model.parameters())
opt = torch.optim.Adam(params, lr=1e-3, weight_decay=5e-4, fused=True)
g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n, dev, is_sorted=True)
res = ResidentLP(g, torch.from_numpy(lp.c_feas).to(dev), torch.from_numpy(lp.v_feas).to(dev), torch.from_numpy(lp.y_s).to(dev), torch.from_numpy(lp.y_t).to(dev))
loader = NeighborSubgraphLoader(res, [6] * 2, 16384, shuffle=True, drop_last=True, seed=1)
state = {'allocs': 0, 'reserved': 0, 'step': 0}
def run(count, stamps=None):
    done = 0
    while done < count:
        for batch in loader:
            batch.to(dev, non_blocking=True)
            lc, lv = model(batch)
            lc, lv = lc[:batch.s_bs], lv[:batch.t_bs]
            loss = balanced(lc, lv, batch.y_s[:batch.s_bs], batch.y_t[:batch.t_bs])
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            done += 1
            ms = torch.cuda.memory_stats()
            if ms["num_device_alloc"] != state["allocs"]:
                print(f"  step {state['step']}: cudaMalloc x{ms['num_device_alloc'] - state['allocs']}, reserved {state['reserved']/2**20:.0f} -> {ms['reserved_bytes.all.current']/2**20:.0f} MiB "
                      f"(batch: {batch.x_s.shape[0]} + {batch.x_t.shape[0]} nodes, nnz {batch.edge_index.nnz()})", flush=True)
                state["allocs"], state["reserved"] = ms["num_device_alloc"], ms["reserved_bytes.all.current"]
            state["step"] += 1
            if stamps is not None: stamps.append(time.perf_counter())
            if done >= count: break
run(10)
torch.cuda.synchronize()
for rep in range(6):
    st = []
    torch.cuda.synchronize(); t0 = time.perf_counter()
    run(30, st)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    d = np.diff([t0] + st) * 1e3
    print(f"rep {rep}: {(t1-t0)/30*1e3:.3f} ms/step; host per-step max {d.max():.2f} median {np.median(d):.2f}; reserved {torch.cuda.memory_reserved()/2**30:.1f} GiB, num_alloc_retries {torch.cuda.memory_stats()['num_alloc_retries']}, segments {torch.cuda.memory_stats()['segment.all.allocated']}", flush=True)
