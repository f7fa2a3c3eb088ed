"""Is the C2 step power-bound?  Runs the fp16 prediction step in a loop for a few seconds and samples nvidia-smi
(power.draw, enforced.power.limit, clocks.sm, clocks_throttle_reasons.sw_power_cap) every 50 ms next to it; then the same
for the hidden transform pair alone and the SpMM pair alone."""
import os, subprocess, sys, threading, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200  # noqa: F401
from lpgnn_b200 import arch, ops, synth
from lpgnn_b200.graph import BipartiteCSR

dev = torch.device("cuda:0")
cfg = synth.CONFIGS["C2"]
lp = synth.processed_lp(cfg["m"], cfg["n"], cfg["nnz"], seed=cfg["seed"])
torch.manual_seed(0)
model = arch.GCN_FC(8, 8, hids=1024, depth=3).to(dev).eval().set_precision("fp16")
t = lambda a, dt: torch.from_numpy(a.astype(dt)).to(dev)
row, col, val = t(lp.row, np.int32), t(lp.col, np.int32), t(lp.a_data, np.float32)
xs, xt = torch.from_numpy(lp.c_feas).to(dev), torch.from_numpy(lp.v_feas).to(dev)
g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n, dev, is_sorted=True)
csr, csc = g.views()
L = torch.randn(lp.m, 1024, device=dev).half(); R = torch.randn(lp.n, 1024, device=dev).half()
conv = model.layers[0]
w16 = lambda p: p.detach().half().contiguous()
Wl = (w16(conv.left2right.lin_rel.weight), w16(conv.left2right.lin_root.weight), conv.left2right.lin_rel.bias.detach())
agg_t = torch.randn(lp.n, 1024, device=dev).half()


def sample(stop, out):
    q = "power.draw,enforced.power.limit,clocks.sm,clocks_throttle_reasons.sw_power_cap"
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "-i", "0", f"--query-gpu={q}", "--format=csv,noheader,nounits"], capture_output=True, text=True)
        try:
            p, lim, clk, cap = [x.strip() for x in r.stdout.strip().split(",")]
            out.append((float(p), float(lim), float(clk), cap))
        except Exception:
            pass
        time.sleep(0.05)


def run(name, f, seconds=4.0):
    for _ in range(20): f()
    torch.cuda.synchronize()
    stop, out = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, out)); th.start()
    n, t0 = 0, time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(50): f()
        n += 50
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); th.join()
    out = out[len(out) // 4:]
    ms = e0.elapsed_time(e1) / n
    print(f"{name:34s} {ms*1e3:8.1f} us/iter   power {np.mean([o[0] for o in out]):6.0f} W of {out[0][1]:.0f} W cap   "
          f"sm clock {np.median([o[2] for o in out]):5.0f} MHz   sw_power_cap active in {np.mean([o[3] == 'Active' for o in out])*100:3.0f} % of samples", flush=True)


with torch.no_grad():
    run("C2 fp16 prediction step", lambda: model.predict_basis_coo(row, col, val, lp.m, lp.n, xs, xt, is_sorted=True))
    run("hidden transform (vars side) alone", lambda: ops.node_transform(agg_t, Wl[0], R, Wl[1], Wl[2], relu=True))
    run("SpMM pair alone", lambda: (ops.spmm(csc, L), ops.spmm(csr, R)))
    run("input layer pair alone", lambda: ops.conv_in_16_pair(csr, csc, xs, xt,
        (model.conv1.left2right.lin_rel.weight.detach(), model.conv1.left2right.lin_rel.bias.detach(), model.conv1.left2right.lin_root.weight.detach()),
        (model.conv1.right2left.lin_rel.weight.detach(), model.conv1.right2left.lin_rel.bias.detach(), model.conv1.right2left.lin_root.weight.detach()),
        torch.float16))
