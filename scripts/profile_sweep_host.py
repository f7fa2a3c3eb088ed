"""Host-side profile (cProfile) of the packed C5 sweep: where the per-LP Python / driver time goes."""
import cProfile, os, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200  # noqa: F401
from lpgnn_b200 import arch, synth
from lpgnn_b200.pipeline import PackedBasisPipeline, pack_lp
dev = torch.device("cuda:0")
pop = synth.lp_population(96, seed=1239)
lps = [synth.processed_lp(m, n, z, seed=sd) for (m, n, z, sd) in pop]
hosts = [pack_lp(lp.row, lp.col, lp.a_data, lp.c_feas, lp.v_feas, is_sorted=True) for lp in lps]
torch.manual_seed(0)
model = arch.GCN_FC(8, 8, hids=1024, depth=3).to(dev).eval().set_precision("fp16")
pipe = PackedBasisPipeline(model, dev)
seq = [hosts[i % 96] for i in range(2000)]
for _ in pipe.run(seq[:200]): pass
torch.cuda.synchronize()
t0 = time.perf_counter(); n = sum(1 for _ in pipe.run(seq)); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{n / dt:.0f} LPs/s e2e ({dt / n * 1e6:.1f} us per LP), packs: {len(pipe._plan(seq))}")
pr = cProfile.Profile(); pr.enable()
for _ in pipe.run(seq): pass
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
