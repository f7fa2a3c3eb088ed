"""Sampled mini-batch path (train.py:105-116: NeighborLoader fan-out 6 per hop over an LP resident on the device): where a
mini-batch step spends its time -- neighbour sampling + induced-subgraph build (``NeighborSubgraphLoader.sample``) vs
the model's forward + loss + backward on the sampled batch.  Host wall time with a synchronize on both sides (the
sampler has host reads of frontier sizes, so its cost is not visible to CUDA events alone) and CUDA-event time.

    python scripts/bench_sampling.py [seeds_per_batch] [hops] [fanout]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200  # noqa: E402,F401
from lpgnn_b200 import arch, synth  # noqa: E402
from lpgnn_b200.graph import BipartiteCSR  # noqa: E402
from lpgnn_b200.losses import balanced  # noqa: E402
from lpgnn_b200.sampling import NeighborSubgraphLoader, ResidentLP  # noqa: E402

seeds_per_batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16_384
hops = int(sys.argv[2]) if len(sys.argv) > 2 else 2
fanout = int(sys.argv[3]) if len(sys.argv) > 3 else 6
dev = torch.device("cuda:0")
cfg = synth.CONFIGS["C3"]
lp = synth.processed_lp(cfg["m"], cfg["n"], cfg["nnz"], seed=cfg["seed"])
g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n, dev, is_sorted=True)
res = ResidentLP(g, torch.from_numpy(lp.c_feas).to(dev), torch.from_numpy(lp.v_feas).to(dev),
                 torch.from_numpy(lp.y_s).to(dev), torch.from_numpy(lp.y_t).to(dev))
loader = NeighborSubgraphLoader(res, [fanout] * hops, seeds_per_batch, shuffle=True, drop_last=True, seed=1)
torch.manual_seed(0)
model = arch.GCN_FC(8, 8, hids=cfg["hids"], depth=cfg["depth"]).to(dev).train().set_precision("bf16")


def timed(fn, reps):
    """(host ms, device ms) per call."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps, a.elapsed_time(b) / reps


order = torch.randperm(res.num_nodes, device=dev)
state = {"i": 0}


def sample_one():
    i = state["i"] = (state["i"] + 1) % (res.num_nodes // seeds_per_batch)
    return loader.sample(order[i * seeds_per_batch:(i + 1) * seeds_per_batch], salt=i)


batch = sample_one()
nodes, nnz = batch.x_s.shape[0] + batch.x_t.shape[0], batch.edge_index.nnz()


def model_step():
    lc, lv = model(batch)
    loss = balanced(lc[:batch.s_bs], lv[:batch.t_bs], batch.y_s[:batch.s_bs], batch.y_t[:batch.t_bs])
    model.zero_grad(set_to_none=True)
    loss.backward()


h_s, d_s = timed(sample_one, 20)
h_m, d_m = timed(model_step, 20)
print(f"seeds {seeds_per_batch}, fan-out [{fanout}]*{hops}: sampled {nodes} nodes / {nnz} nnz of {res.num_nodes} / {lp.nnz}")
print(f"sample + induced subgraph : {h_s:.3f} ms host, {d_s:.3f} ms device")
print(f"forward + loss + backward : {h_m:.3f} ms host, {d_m:.3f} ms device")
