#!/usr/bin/env python
"""bench.py -- the lp-gnn hot path on B200 (contract: see the task prompt / DESIGN.md section 6).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path (oracle port)

A *step* is one pass of the hot path over one synthetic LP of the named workload (default: C2 of
BASELINE.json, GCN_FC(8,8,hids=1024,depth=3) inference on a 50K x 100K LP with ~500K nonzeros):

    graph build (COO -> CSR+CSC)  ->  GCN_FC forward (conv1, SpMM, node transforms, head+mask)
    ->  basis selection (softmax / top-m / status)

* ``value``  LPs/s with the step's inputs (COO, node features) already resident in HBM.
* ``e2e``    the same metric through the public API (``pipeline.BasisPipeline``) with HOST (pinned) inputs:
             per step one H2D copy of the packed COO + features (prefetched on a side stream while the previous LP
             computes), the step, and a D2H copy of the status vector, all inside the timed region.
* ``roofline``      dominant kernel (tcgen05 node transform) against the measured bf16 peak, timed
                    live with CUDA events; ``kernels`` lists every kernel of the step the same way
                    (HBM-bound ones against the measured copy bandwidth).
* ``cpu_baseline``  the oracle port of the reference's CPU path, timed on this box's host cores on a
                    bounded sample (rank 0, N=1 only).
* ``fp32``          the same step at the reference's DEFAULT precision (`--fp16 0`): fp32 storage, hidden transforms on
                    the tensor cores from x2 operands (csrc/gemm_x2.cu) -- the like-for-like figure against the fp32 CPU arm.
* ``parity``        logits / statuses of every precision against the CPU oracle on the bench LP (N=1).
With N > 1 (torchrun, one rank per GPU) every rank runs its own LPs (independent units, no data-path
collective): weak scaling, value = N * K LPs / max-over-ranks time.  The two workloads that really shard ride in the
same line at every N: ``train_ms_per_step`` / ``train_lps_per_sec`` (C3: data-parallel training step with the NCCL
gradient all-reduce) and ``sweep_lps_per_sec`` / ``sweep_e2e_lps_per_sec`` (C5: packed sweep, LPs dealt over ranks).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------------- utils
def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        """Starts the sampler and waits for its first sample, so that nvidia-smi's start-up (NVML init takes
        driver locks for ~100 ms) does not land inside a timed region."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < 5.0:
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def mark(self):
        """Samples taken from now on belong to the timed region."""
        self.first = len(self.lines)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines[getattr(self, "first", 0):]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        # NCCL_DEBUG is left to the caller (the driver counts ranks from NCCL's INFO lines); the JSON line is the one
        # stdout line that starts with '{"metric"' (or '{"impl"')
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo", rank=rank, world_size=world,
                                device_id=torch.device("cuda", local) if torch.cuda.is_available() else None)
    return rank, world, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()


def max_over_ranks(x, world, dev):
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def workload_spec(name):
    from lpgnn_b200 import synth
    cfg = dict(synth.CONFIGS[name])
    cfg["name"] = name
    return cfg


def workload_string(name, H, D, structure, m, n, z):
    """config.workload, identical in both arms (the driver compares the strings)."""
    return (f"{name}: GCN_FC(8,8,hids={H},depth={D}) inference, synthetic {structure} LP {m}x{n}, nnz={z}, "
            f"one LP per step per GPU")


def sweep_workload_string(n_distinct, nnz_mean, hids):
    return (f"C5: sweep over {n_distinct} distinct synthetic LPs per GPU (m log-uniform 100..20000, n=2m, nnz~5n, "
            f"mean nnz {nnz_mean:.0f}), GCN_FC(8,8,hids={hids},depth=3), one LP per step")


def mp_edges(nnz, depth, fwd_bwd=False):
    # z * 2 directions * (D-1) conv layers * (1 fwd | 2 fwd+bwd)
    return nnz * 2 * (depth - 1) * (2 if fwd_bwd else 1)


# ----------------------------------------------------------------------------------------------- GPU arm
def time_kernel(fn, reps, flush=None):
    """Average duration (ms) of ``fn`` over ``reps`` launches, CUDA events on the current stream."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    time.sleep(0.05)                           # kernels are timed ALONE (burst conditions, like the burst peaks they are held against)
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()                      # > L2-sized write between timed launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def run_gpu(args):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib, arch, synth

    rank, world, local = dist_setup(args.gpus)
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    peaks = load_peaks()
    cfg = workload_spec(args.workload)
    bf16 = args.precision in ("bf16", "fp16")      # 16-bit storage on the tensor-core path
    H, D = cfg["hids"], cfg["depth"]

    # every rank draws its own ring of distinct LPs of the workload's shape (independent units); the timed loops cycle
    # through the ring, so no step finds its COO / features left in L2 by the step before it
    ring = max(1, args.distinct if cfg["nnz"] <= 2_000_000 else 1)
    lps = [synth.processed_lp(cfg["m"], cfg["n"], cfg["nnz"], seed=cfg["seed"] + 1000 * rank + 17 * i, structure=args.structure)
           for i in range(ring)]
    lp = lps[0]
    m, n, z = lp.m, lp.n, lp.nnz
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=H, depth=D).to(dev).eval()
    model.set_precision(args.precision)

    # host (pinned) inputs of one step: one packed staging buffer [row | col | val | x_s | x_t] (4-byte words)
    from lpgnn_b200.pipeline import BasisPipeline, pack_lp, unpack_device
    hosts = [pack_lp(x.row, x.col, x.a_data, x.c_feas, x.v_feas, is_sorted=True) for x in lps]
    h2d_bytes = float(np.mean([h.nbytes for h in hosts]))
    d2h_bytes = float(np.mean([h.m + h.n for h in hosts]))
    # device-resident copies for the HBM-resident arm
    dev_lps = [(h, tuple(t.clone() for t in unpack_device(h.pack.to(dev), h))) for h in hosts]

    # LPs of a sweep are independent: `--inflight k` enqueues consecutive LPs on alternating streams, so the
    # latency-bound small kernels of one LP (graph build, basis selection) overlap the other LPs' work.
    streams = [torch.cuda.Stream(device=dev) for _ in range(max(args.inflight, 1))] if args.inflight > 1 else None
    step_no = [0]

    def step_resident():
        # the processed-file COO is row-major sorted (dataset.py:208-210: A.tocoo() of a CSR) -> is_sorted hint
        # one native call: graph build + forward + basis selection (lpgnn_predict_basis)
        i = step_no[0]
        step_no[0] += 1
        h, (row, col, val, xs, xt) = dev_lps[i % ring]
        if streams is None:
            return h, model.predict_basis_coo(row, col, val, h.m, h.n, xs, xt, is_sorted=True)
        with torch.cuda.stream(streams[i % len(streams)]):
            return h, model.predict_basis_coo(row, col, val, h.m, h.n, xs, xt, is_sorted=True)

    pipe = BasisPipeline(model, dev, compute_streams=max(args.inflight, 1))

    def run_e2e(k):
        """k steps through the public pipeline API: per step one H2D copy of the packed LP from pinned host memory
        (prefetched on a side stream while the previous LP computes) and one D2H copy of its statuses."""
        ok = True
        for j, st in pipe.run([hosts[i % ring] for i in range(k)]):
            if j == k - 1:
                ok = int((st == 1).sum()) == hosts[j % ring].m
        return ok

    def timed_resident(k):
        """EXACTLY k steps between barrier + synchronize, CUDA events on the launching stream, max over ranks."""
        barrier(world)
        torch.cuda.synchronize()
        l0 = lib.lpgnn_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if streams is not None:
            for st in streams:
                st.wait_stream(torch.cuda.current_stream())
        for _ in range(k):
            h, status = step_resident()
        if streams is not None:
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        launches = lib.lpgnn_launch_count() - l0
        barrier(world)
        n_basic = int((status == 1).sum().item())
        assert n_basic == h.m, f"basis invariant violated: {n_basic} basic nodes for m={h.m}"
        return max_over_ranks(e0.elapsed_time(e1), world, dev), int(launches)

    def timed_e2e(k):
        assert run_e2e(3)
        barrier(world)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        assert run_e2e(k)
        torch.cuda.synchronize()
        t = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
        barrier(world)
        return t

    def measure(precision, k):
        model.set_precision(precision)
        for _ in range(max(args.warmup, 3)):
            step_resident()
        torch.cuda.synchronize()
        t_ms, launches = timed_resident(k)
        t_e2e = timed_e2e(k)
        return t_ms, t_e2e, launches

    # ---- warm-up, then the timed regions of the headline precision
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.mark()
    t_ms, t_e2e, launches = measure(args.precision, args.steps)
    clocks = sampler.stop() if rank == 0 else None

    lps_s = world * args.steps / (t_ms / 1e3)
    lps_e2e = world * args.steps / (t_e2e / 1e3)
    out = {
        "metric": "LPs/sec (basis prediction: graph build + GCN_FC forward + basis selection)",
        "value": lps_s, "unit": "LPs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": t_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.precision if bf16 else "f32", "data": "synthetic",
        "config": {"workload": workload_string(cfg["name"], H, D, args.structure, m, n, z),
                   "ring": f"{ring} distinct LPs of this shape per GPU, cycled (sizes of the first one in `workload`)",
                   "l2": "activations per layer (>=300 MB) exceed the 126 MB L2 and consecutive steps use different LPs; "
                         "no explicit flush in the step loop",
                   "precision": args.precision, "structure": args.structure,
                   "in_flight": f"{max(args.inflight, 1)} LP(s) in flight on alternating streams in both arms "
                                "(LPs are independent units; ms_per_step = 1 / throughput)"},
        "mp_edges_per_sec": mp_edges(z, D) * lps_s,
        "e2e": {"value": lps_e2e, "unit": "LPs/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": t_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }

    # ---- the other precisions on the same workload and protocol (all ranks take part: same barriers)
    if H % 64 == 0 and D > 2 and not args.no_precisions:
        k2 = max(10, min(args.steps, 100))
        others = [p for p in ("fp32", "fp16", "bf16") if p != args.precision]
        for p in others:
            t2, t2e, l2 = measure(p, k2)
            if rank == 0:
                out[p] = {"value": world * k2 / (t2 / 1e3), "unit": "LPs/s", "ms_per_step": t2 / k2, "steps": k2,
                          "e2e": {"value": world * k2 / (t2e / 1e3), "unit": "LPs/s", "ms_per_step": t2e / k2},
                          "gpu_launches": int(l2),
                          "note": f"same workload, ring and timed regions as `value` / `e2e`, precision='{p}'"
                                  + ("; the reference's default arithmetic (--fp16 0): the like-for-like figure against the "
                                     "fp32 CPU arm" if p == "fp32" else "")}
        model.set_precision(args.precision)
    if rank == 0 and world == 1 and not args.no_cpu and H % 64 == 0 and D > 2:
        out["parity"] = parity_vs_oracle(model, cfg, lp, dev_lps[0], args.precision)
    elif rank == 0 and bf16 and H % 64 == 0 and D > 2:
        # cheap live gate when the CPU oracle is not run: statuses against this library's own fp32 mode
        h, (row, col, val, xs, xt) = dev_lps[0]
        st16 = model.predict_basis_coo(row, col, val, h.m, h.n, xs, xt, is_sorted=True)
        model.set_precision("fp32")
        st32 = model.predict_basis_coo(row, col, val, h.m, h.n, xs, xt, is_sorted=True)
        model.set_precision(args.precision)
        out["status_agreement_vs_gpu_fp32_mode"] = float((st32 == st16).float().mean().item())

    # ---- the workloads that shard: C3 data-parallel training step, C5 packed sweep (every N, same line)
    if not args.no_train:
        tp = "bf16" if args.precision == "fp16" else ("fp32" if args.precision.startswith("fp32") else args.precision)
        tr = train_throughput(cfg, lp, dev, tp, args.train_steps or max(10, min(args.steps, 50)), 3, world)
        c1 = workload_spec("C1")
        lp1 = synth.processed_lp(c1["m"], c1["n"], c1["nnz"], seed=c1["seed"] + 1000 * rank, structure=args.structure)
        tr1 = train_throughput(c1, lp1, dev, "fp32", max(10, min(args.steps, 100)), 3, world)
        tr1p = train_packed_throughput(c1, dev, "fp32", 64, max(10, min(args.steps, 50)), world, rank, args.structure)
        ts = None
        if args.workload == "C3":
            try:
                ts = train_sampled_throughput(cfg, lp, dev, tp, max(10, min(args.steps, 30)), world)
            except Exception as e:          # keep the bench line: the sampled variant is an extra of --workload C3
                if world > 1:
                    raise                   # ranks must not diverge around collectives
                ts = {"error": f"{type(e).__name__}: {e}"}
        if rank == 0:
            out["train"] = tr
            out["train_ms_per_step"] = tr["ms_per_step"]
            out["train_lps_per_sec"] = tr["lps_per_sec"]
            out["train_c1_fp32"] = tr1
            out["train_c1_fp32_packed"] = tr1p
            if ts is not None:
                out["train_sampled"] = ts
            if world == 1 and not args.no_cpu:
                out["train_c1_fp32"]["cpu_baseline"] = cpu_train_baseline(c1, lp1)
                out["train"]["cpu_baseline"] = cpu_train_baseline(cfg, lp, sample_budget_s=20.0, max_steps=3, warm=1)
    if not args.no_sweep and args.workload in ("C2", "C3"):
        # (at least 1000 LPs = ~20 packs per rank: a handful of packs would only time the pipeline's fill and drain)
        sw = sweep_measure(args, rank, world, dev, lib, args.precision, steps=max(1000, min(10 * args.steps, 3000)),
                           distinct=args.sweep_distinct, with_cpu=False)
        sw.pop("_model"), sw.pop("_largest")
        if rank == 0:
            out["sweep"] = sw
            out["sweep_lps_per_sec"] = sw["value"]
            out["sweep_e2e_lps_per_sec"] = sw["e2e"]["value"]
    if rank == 0 and not args.no_kernels:
        out.update(kernel_rooflines(model, lp, dev, peaks, bf16, args))
    if rank == 0 and world == 1 and not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(cfg, lp, args, sample_budget_s=args.cpu_seconds)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


# Known deviations of the 16-bit storage modes from the north-star logit bar (2e-2 of the row norm on EVERY entry): with
# random-initialised weights a handful of rows have a raw logit vector 10-100x shorter than typical and F.normalize
# (arch.py:134-135) amplifies their rounding error by that factor.  Stated in the record, asserted in the tests.
KNOWN_DEVIATION = {
    "bf16": "worst entry ~2e-1 (bar 2e-2) and status agreement ~99.8 % (bar 99.9 %) at C2 size: bf16's 8 mantissa bits; "
            "the gate-passing 16-bit mode is fp16",
    "fp16": "a handful of entries reach ~4e-2 (bar 2e-2) at C2 size; status agreement >= 99.9 % holds",
}


def parity_vs_oracle(model, cfg, lp, dev_lp, headline):
    """Logits and statuses of every precision against the CPU oracle port (fp32, the reference's arithmetic) on the
    bench LP: max entry error relative to the row norm 10 that add_knowledge imposes, share of entries within the
    north star's 16-bit bar 2e-2, status agreement with val.inference_gnn on the oracle's logits."""
    port, ref, g = _port_setup(cfg, lp)
    with torch.no_grad():
        ec, ev = ref(torch.from_numpy(lp.c_feas), torch.from_numpy(lp.v_feas), port.TorchGraph(g))
    exp = torch.cat((ec, ev)).numpy()
    exp_status = port.inference_gnn_np(exp, lp.m)
    h, (row, col, val, xs, xt) = dev_lp
    res = {}
    for p in ("fp32", "fp16", "bf16"):
        model.set_precision(p)
        st, lg = model.predict_basis_coo(row, col, val, h.m, h.n, xs, xt, is_sorted=True, want_logits=True)
        d = np.abs(lg.cpu().numpy() - exp) / 10.0
        res[p] = {"max_err": float(d.max()), "frac_within_2e-2": float(np.mean(d < 2e-2)),
                  "status_agreement_vs_oracle": float(np.mean(st.cpu().numpy() == exp_status)),
                  "bar": "max_err < 1e-4" if p == "fp32" else "max_err < 2e-2, status agreement >= 0.999"}
        if p in KNOWN_DEVIATION:
            res[p]["known_deviation"] = KNOWN_DEVIATION[p]
    model.set_precision(headline)
    res["oracle"] = "oracle/port.py (CPU, fp32; the reference's arch.py / val.py arithmetic) on ring LP 0, same weights"
    return res


def sweep_measure(args, rank, world, dev, lib, precision, steps, distinct, with_cpu):
    """The batch basis-prediction sweep (scripts/pred_basis.py workload, BASELINE config C5) over a population of
    small/medium LPs (m log-uniform in [100, 20000], n = 2m, nnz = 5n).  A step = one LP.  The LPs are
    independent units: with N ranks the population is N times larger and dealt over the ranks (no collective)."""
    from lpgnn_b200 import arch, synth
    from lpgnn_b200.pipeline import BasisPipeline, PackedBasisPipeline, pack_lp, unpack_device
    # weak scaling: `distinct` LPs PER RANK.  The population of world * distinct sizes is dealt in descending-nnz
    # order, forwards then backwards over the ranks ("snake"), so every rank holds the same number of LPs and a
    # near-equal share of the work -- a rank-dependent size mix would only measure the deal.
    pop = synth.lp_population(distinct * world, seed=1239)
    if world > 1:
        from lpgnn_b200.io_utils import shard_indices
        mine = [pop[i] for i in shard_indices(len(pop), rank, world, weights=[p[2] for p in pop], equal_counts=True)]
        perm = np.random.default_rng(1239 + rank).permutation(len(mine))     # sweep order: shuffled, not sorted
        mine = [mine[i] for i in perm]
    else:
        mine = pop
    lps = [synth.processed_lp(m, n, z, seed=sd) for (m, n, z, sd) in mine]
    hosts = [pack_lp(lp.row, lp.col, lp.a_data, lp.c_feas, lp.v_feas, is_sorted=True) for lp in lps]
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=args.sweep_hids, depth=3).to(dev).eval().set_precision(precision)
    dev_lps = [(h, tuple(t.clone() for t in unpack_device(h.pack.to(dev), h))) for h in hosts]

    def one(i):
        h, (row, col, val, xs, xt) = dev_lps[i % len(dev_lps)]
        return model.predict_basis_coo(row, col, val, h.m, h.n, xs, xt, is_sorted=True)

    for i in range(max(args.warmup, 3)):
        one(i)
    barrier(world)
    torch.cuda.synchronize()
    l0 = lib.lpgnn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(i)
    e1.record()
    torch.cuda.synchronize()
    launches = lib.lpgnn_launch_count() - l0
    barrier(world)
    t_ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    pipe = PackedBasisPipeline(model, dev) if args.sweep_pack else BasisPipeline(model, dev)
    seq = [hosts[i % len(hosts)] for i in range(steps)]
    for _ in pipe.run(seq[:min(len(seq), 3 * len(hosts))]):      # warm-up: full-size packs, so the grow-only buffers exist
        pass
    barrier(world)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n_done = sum(1 for _ in pipe.run(seq))
    torch.cuda.synchronize()
    t_e2e = max_over_ranks((time.perf_counter() - t0) * 1e3, world, dev)
    barrier(world)
    assert n_done == steps
    nnz_mean = float(np.mean([lp.nnz for lp in lps]))
    lps_s, lps_e2e = world * steps / (t_ms / 1e3), world * steps / (t_e2e / 1e3)
    out = {
        "value": lps_s, "unit": "LPs/s", "steps": steps, "ms_per_step": t_ms / steps, "scaling": "weak",
        "workload": sweep_workload_string(len(mine), nnz_mean, args.sweep_hids),
        "deal": "population of N x this many sizes dealt over the ranks in descending-nnz snake order (equal counts, "
                "near-equal work), shuffled within the rank; no collective on the data path",
        "precision": precision,
        "e2e_api": "PackedBasisPipeline (block-diagonal packs, segmented basis decision)" if args.sweep_pack
        else "BasisPipeline (one native call per LP)",
        "mp_edges_per_sec": mp_edges(nnz_mean, 3) * lps_s,
        "e2e": {"value": lps_e2e, "unit": "LPs/s", "h2d_bytes_per_step": float(np.mean([h.nbytes for h in hosts])),
                "d2h_bytes_per_step": float(np.mean([h.m + h.n for h in hosts])), "ms_per_step": t_e2e / steps},
        "gpu_launches": int(launches)}
    if with_cpu and rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_sweep_baseline(lps, args.sweep_hids, sample_budget_s=args.cpu_seconds)
    out["_model"], out["_largest"] = model, max(lps, key=lambda lp: lp.nnz)      # for the caller's kernel table; popped there
    return out


def run_sweep(args):
    """--workload C5: the sweep as the headline line (see sweep_measure)."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib
    rank, world, local = dist_setup(args.gpus)
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.mark()
    sw = sweep_measure(args, rank, world, dev, lib, args.precision, args.steps, args.sweep_distinct, not args.no_cpu)
    clocks = sampler.stop() if rank == 0 else None
    model, largest = sw.pop("_model"), sw.pop("_largest")
    roof = None
    if rank == 0 and not args.no_kernels:
        # per-kernel table on the LARGEST LP of this rank's share (every kernel timed alone, L2 flushed): the sweep has no
        # single launch shape, and its packs (600K nodes) run the same kernels at C2-like sizes
        roof = kernel_rooflines(model, largest, dev, load_peaks(), args.precision in ("bf16", "fp16"), args)
        roof["roofline"]["workload"] = f"largest LP of the population ({largest.m} x {largest.n}, nnz {largest.nnz}), one LP per launch"
    if rank == 0:
        out = {
            "metric": "LPs/sec (basis prediction: graph build + GCN_FC forward + basis selection)",
            "value": sw["value"], "unit": "LPs/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": sw["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.precision if args.precision in ("bf16", "fp16") else "f32", "data": "synthetic",
            "config": {"workload": sw["workload"], "deal": sw["deal"], "precision": args.precision,
                       "l2": "small LPs: working set is L2-resident by nature of the workload", "e2e_api": sw["e2e_api"]},
            "mp_edges_per_sec": sw["mp_edges_per_sec"], "e2e": sw["e2e"], "gpu_launches": sw["gpu_launches"],
            "clocks": clocks}
        if "cpu_baseline" in sw:
            out["cpu_baseline"] = sw["cpu_baseline"]
        if roof is not None:
            out.update(roof)
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def train_throughput(cfg, lp, dev, precision, steps, warmup, world):
    """Training step of the same shape (reference train.py:117-129): forward, balanced loss, backward through the
    CUDA kernels, gradient all-reduce (N > 1) and Adam, one LP graph per step per GPU.  Returns steps/s and
    message-passing edges/s (fwd+bwd): nnz * 2 directions * (depth-1) conv layers * 2."""
    import types as _t
    from lpgnn_b200 import arch
    from lpgnn_b200.graph import BipartiteCSR
    from lpgnn_b200.losses import balanced
    from lpgnn_b200.train import allreduce_gradients, broadcast_parameters
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=cfg["hids"], depth=cfg["depth"]).to(dev).train().set_precision(precision)
    broadcast_parameters(model, world)
    params = list(model.parameters())
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=5e-4, fused=True)
    g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n, dev, is_sorted=True)
    batch = _t.SimpleNamespace(x_s=torch.from_numpy(lp.c_feas).to(dev), x_t=torch.from_numpy(lp.v_feas).to(dev), edge_index=g)
    y_s, y_t = torch.from_numpy(lp.y_s).to(dev), torch.from_numpy(lp.y_t).to(dev)

    def step():
        lc, lv = model(batch)
        loss = balanced(lc, lv, y_s, y_t)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        allreduce_gradients(params, world)
        opt.step()
        return loss

    for _ in range(max(warmup, 3)):
        step()
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    t_ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    assert bool(torch.isfinite(loss))
    sps = world * steps / (t_ms / 1e3)
    return {"workload": f"{cfg['name']}-shaped LP, GCN_FC(8,8,hids={cfg['hids']},depth={cfg['depth']}) training step "
                        f"(fwd + balanced loss + bwd + Adam{' + NCCL grad all-reduce' if world > 1 else ''}), dp={model.dp}",
            "precision": precision, "steps": steps, "ms_per_step": t_ms / steps, "lps_per_sec": sps,
            "mp_edges_per_sec_fwd_bwd": mp_edges(lp.nnz, cfg["depth"], fwd_bwd=True) * sps}


def train_packed_throughput(cfg, dev, precision, n_lps, steps, world, rank, structure):
    """Mini-batches of LP graphs (north star config 3; `train.py --pack`): ``n_lps`` distinct LPs of the workload's shape
    packed block-diagonally, ONE native forward / backward over the pack, the per-LP balanced loss averaged over the pack
    (losses.balanced_packed), gradient all-reduce (N > 1), Adam.  Reported per LP: the reference's loop (train.py:70,
    117-129) takes one optimiser step per LP, here ``n_lps`` graphs share one."""
    from lpgnn_b200 import arch, dataset, synth
    from lpgnn_b200.data import Data
    from lpgnn_b200.graph import BipartiteCSR
    from lpgnn_b200.losses import balanced_packed
    from lpgnn_b200.train import allreduce_gradients, broadcast_parameters
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=cfg["hids"], depth=cfg["depth"]).to(dev).train().set_precision(precision)
    broadcast_parameters(model, world)
    params = list(model.parameters())
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=5e-4, fused=True)
    items = []
    for i in range(n_lps):
        lp = synth.processed_lp(cfg["m"], cfg["n"], cfg["nnz"], seed=cfg["seed"] + 1000 * rank + 31 * i, structure=structure)
        g = BipartiteCSR.from_coo(torch.from_numpy(lp.row), torch.from_numpy(lp.col), torch.from_numpy(lp.a_data.astype(np.float32)),
                                  lp.m, lp.n, is_sorted=True)
        items.append(Data(x_s=torch.from_numpy(lp.c_feas), x_t=torch.from_numpy(lp.v_feas), y_s=torch.from_numpy(lp.y_s),
                          y_t=torch.from_numpy(lp.y_t), edge_index=g))
    batch = dataset.pack_bipartite(items).to(dev)
    nnz = batch.edge_index.nnz()

    def step():
        lc, lv = model(batch)
        loss = balanced_packed(lc, lv, batch.y_s, batch.y_t, batch.cons_ptr, batch.vars_ptr)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        allreduce_gradients(params, world)
        opt.step()
        return loss

    for _ in range(3):
        step()
    barrier(world)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    barrier(world)
    t_ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    assert bool(torch.isfinite(loss))
    lps = world * steps * n_lps / (t_ms / 1e3)
    return {"workload": f"packs of {n_lps} {cfg['name']}-shaped LPs ({batch.x_s.shape[0]} x {batch.x_t.shape[0]}, nnz {nnz} per pack), "
                        f"GCN_FC(8,8,hids={cfg['hids']},depth={cfg['depth']}): one training step per pack (fwd + per-LP balanced "
                        f"loss + bwd + Adam{' + NCCL grad all-reduce' if world > 1 else ''})",
            "precision": precision, "steps": steps, "ms_per_step": t_ms / steps, "ms_per_lp": t_ms / steps / n_lps,
            "lps_per_sec": lps, "mp_edges_per_sec_fwd_bwd": mp_edges(nnz / n_lps, cfg["depth"], fwd_bwd=True) * lps}


def kernel_rooflines(model, lp, dev, peaks, bf16, args):
    """Per-kernel durations (CUDA events, live) and roofline fractions with the algorithmic bytes /
    flops of SURVEY 8d (stated in DESIGN.md section 4)."""
    from lpgnn_b200 import ops
    from lpgnn_b200.graph import BipartiteCSR
    m, n, z = lp.m, lp.n, lp.nnz
    H = model.hids
    torch.cuda.synchronize()
    time.sleep(1.0)          # the arms before this one ran the GPU at its power cap for seconds: let the boost state recover
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(args.precision, torch.float32)
    s = 2 if bf16 else 4
    g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), m, n, dev)
    csr, csc = g.views()
    xs, xt = torch.from_numpy(lp.c_feas).to(dev), torch.from_numpy(lp.v_feas).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    reps = args.kernel_reps
    kernels = []

    def add(name, bound, t_ms, alg, launches_per_step):
        if bound == "hbm":
            ach = alg / (t_ms * 1e-3) / 1e9
            peak, unit = peaks["hbm_gbs"], "GB/s"
        else:
            ach = alg / (t_ms * 1e-3) / 1e12
            peak, unit = peaks["bf16_burst"], "TFLOP/s"
        kernels.append({"kernel": name, "bound": bound, "ms": t_ms, "algorithmic": alg, "achieved": ach, "peak": peak,
                        "unit": unit, "frac": ach / peak, "launches_per_step": launches_per_step})

    c1 = model.conv1
    w = lambda p: p.detach()
    # input layer, variables side (rows = n): gather [A^T x_s | x_t] then the transform
    if bf16:
        # input layer, ONE kernel for both directions (aggregate + 16-wide MMA + bias + ReLU + 16-bit store); HBM-bound on
        # the output write (a pure write stream of this size runs at ~5.8 TB/s on this part, scripts/bench_store.cu)
        l2r_w = (w(c1.left2right.lin_rel.weight), w(c1.left2right.lin_rel.bias), w(c1.left2right.lin_root.weight))
        r2l_w = (w(c1.right2left.lin_rel.weight), w(c1.right2left.lin_rel.bias), w(c1.right2left.lin_root.weight))
        f_pair = lambda: ops.conv_in_16_pair(csr, csc, xs, xt, l2r_w, r2l_w, dt, relu=True)
        t = time_kernel(f_pair, reps, flush)
        add("conv_in_16_pair (input layer: gather + MMA + ReLU + store, both sides, one launch)", "hbm", t,
            (m + n) * H * s + 2 * (m + n) * 8 * 4 + 2 * z * 8 + (m + n + 2) * 4, 1)
        left, right = f_pair()[:2]
    elif args.precision == "fp32" and len(model.layers):
        # fp32 on the tensor cores: the input layer writes fp32 (gather source of the aggregation) + x2 operands (hi / lo halves)
        f_in = lambda: ops.conv_in_fused_x2(csc, xs, xt, w(c1.left2right.lin_rel.weight), w(c1.left2right.lin_rel.bias),
                                            w(c1.left2right.lin_root.weight), relu=True)
        f_in_s = lambda: ops.conv_in_fused_x2(csr, xt, xs, w(c1.right2left.lin_rel.weight), w(c1.right2left.lin_rel.bias),
                                              w(c1.right2left.lin_root.weight), relu=True)
        t = time_kernel(f_in, reps, flush) + time_kernel(f_in_s, reps, flush)
        add("conv_in_fused_x2 pair (gather + CUDA-core transform, fp32 + x2 outputs, both sides)", "hbm", t,
            (m + n) * H * 8 + 2 * (m + n) * 8 * 4 + 2 * z * 8 + (m + n + 2) * 4, 6)
        right, x2_t, sx_t = f_in()
        left, x2_s, sx_s = f_in_s()
    else:
        f_in = lambda: ops.conv_in_fused(csc, xs, xt, w(c1.left2right.lin_rel.weight), w(c1.left2right.lin_rel.bias),
                                         w(c1.left2right.lin_root.weight), dt, relu=True)
        t = time_kernel(f_in, reps, flush)
        add("conv_in_fused (gather + CUDA-core transform, vars side)", "hbm", t, n * H * s + (m + n) * 8 * 4 + z * 8 + (n + 1) * 4, 2)
        right, _ = f_in()
        left, _ = ops.conv_in_fused(csr, xt, xs, w(c1.right2left.lin_rel.weight), w(c1.right2left.lin_rel.bias),
                                    w(c1.right2left.lin_root.weight), dt, relu=True)
    if len(model.layers):
        conv = model.layers[-1]
        cast = conv._cache.get
        # SpMM pair of one hidden layer: compulsory bytes 2(m+n)Hs + 16z + 4(m+n+2)   
        if args.precision == "fp32":   # fp32 features in, x2 operands out (same bytes per element: 2 + 2)
            t_p = time_kernel(lambda: ops.spmm_x2_pair(csr, csc, left, right, sx_s, sx_t, nnz=z), reps, flush)
            add("spmm_x2 pair (A.R and A^T.L, fp32 in, x2 operands out)", "hbm", t_p,
                2 * (m + n) * H * s + 16 * z + 4 * (m + n + 2) + 8 * (m + n), 1)
        else:   # both aggregations in one launch (as the step runs them)
            t_p = time_kernel(lambda: ops.spmm_pair(csr, csc, left, right, nnz=z), reps, flush)
            add("spmm pair (A.R and A^T.L)", "hbm", t_p, 2 * (m + n) * H * s + 16 * z + 4 * (m + n + 2), 1)
        kernels[-1]["gather_model_bytes"] = 2 * z * H * s + (m + n) * H * s + 16 * z
        agg_t, agg_s = ops.spmm(csc, left), ops.spmm(csr, right)
        l2r, r2l = conv.left2right, conv.right2left
        hl = (w(model.lin_left.weight), w(model.lin_left.bias), xs)
        hr = (w(model.lin_right.weight), w(model.lin_right.bias), xt)
        if bf16:
            # the last hidden layer as the step runs it: basis-status head fused into the epilogue (the activation is
            # never written) + head_finish (sum of the column-tile partials, bias, add_knowledge)
            f_t = lambda: ops.node_transform_head(agg_t, cast(l2r.lin_rel.weight, dt), right, cast(l2r.lin_root.weight, dt),
                                                  w(l2r.lin_rel.bias), *hr)
            f_s = lambda: ops.node_transform_head(agg_s, cast(r2l.lin_rel.weight, dt), left, cast(r2l.lin_root.weight, dt),
                                                  w(r2l.lin_rel.bias), *hl)
            t_g = time_kernel(f_t, reps, flush) + time_kernel(f_s, reps, flush)
            add("node_transform pair, head fused (tcgen05) + head_finish", "tensor", t_g, 4 * (m + n) * H * H, 4)
        elif args.precision == "fp32":
            from lpgnn_b200.autograd import x2_weights_cached
            (at, st_), (as_, ss_) = ops.spmm_x2(csc, left, sx_s), ops.spmm_x2(csr, right, sx_t)
            xt2, xs2 = x2_t, x2_s
            wr_t, wo_t, cs_t = x2_weights_cached(conv._cache, l2r)
            wr_s, wo_s, cs_s = x2_weights_cached(conv._cache, r2l)
            f_t = lambda: ops.node_transform_x2(at, wr_t, xt2, wo_t, st_, cs_t, w(l2r.lin_rel.bias), relu=True, head=hr,
                                                want_out=False, rowscale2=sx_t)
            f_s = lambda: ops.node_transform_x2(as_, wr_s, xs2, wo_s, ss_, cs_s, w(r2l.lin_rel.bias), relu=True, head=hl,
                                                want_out=False, rowscale2=sx_s)
            t_g = time_kernel(f_t, reps, flush) + time_kernel(f_s, reps, flush)
            # three half x half passes carry one fp32-accurate product: 3 x 4(m+n)H^2 tensor-core flops
            add("node_transform_x2 pair, head fused (tcgen05, 3 half passes) + head_finish", "tensor", t_g,
                3 * 4 * (m + n) * H * H, 4)
            kernels[-1]["fp32_equivalent_flops"] = 4 * (m + n) * H * H
        else:
            f_t = lambda: ops.node_transform(agg_t, cast(l2r.lin_rel.weight, dt), right, cast(l2r.lin_root.weight, dt),
                                             w(l2r.lin_rel.bias), relu=True)
            f_s = lambda: ops.node_transform(agg_s, cast(r2l.lin_rel.weight, dt), left, cast(r2l.lin_root.weight, dt),
                                             w(r2l.lin_rel.bias), relu=True)
            t_g = time_kernel(f_t, reps, flush) + time_kernel(f_s, reps, flush)
            add("node_transform pair (fp32 CUDA cores)", "tensor", t_g, 4 * (m + n) * H * H, 2)
            right2 = f_t()
            t_h = time_kernel(lambda: ops.head_mask(right2, w(model.lin_right.weight), w(model.lin_right.bias), xt), reps, flush)
            add("head_mask (vars side)", "hbm", t_h, n * (H * s + 8 * 4 + 3 * 4), 1)
    else:
        t_h = time_kernel(lambda: ops.head_mask(right, w(model.lin_right.weight), w(model.lin_right.bias), xt), reps, flush)
        add("head_mask (vars side)", "hbm", t_h, n * (H * s + 8 * 4 + 3 * 4), 1)
    lc = torch.randn(m, 3, device=dev)
    lv = torch.randn(n, 3, device=dev)
    t_sel = time_kernel(lambda: ops.basis_select(lc, lv, int64=False), reps, flush)
    add("basis_select (one cooperative launch)", "hbm", t_sel, (m + n) * (12 + 1), 1)
    h_row = torch.from_numpy(lp.row.astype(np.int32)).to(dev)
    h_col = torch.from_numpy(lp.col.astype(np.int32)).to(dev)
    h_val = torch.from_numpy(lp.a_data.astype(np.float32)).to(dev)
    t_b = time_kernel(lambda: BipartiteCSR.from_coo(h_row, h_col, h_val, m, n, is_sorted=True), reps, flush)
    add("graph_build (sorted COO->CSR+CSC)", "hbm", t_b, z * 12 * 2 * 2, -1)
    t_b2 = time_kernel(lambda: BipartiteCSR.from_coo(h_row, h_col, h_val, m, n, is_sorted=False), reps, flush)
    add("graph_build (unsorted COO->CSR+CSC)", "hbm", t_b2, z * 12 * 2 * 3, -1)
    # measured DRAM traffic per launch from the committed ncu --set full capture (same workload / precision only)
    for tname in ("r02e_traffic.json", "r01d_traffic.json", "r01e_traffic.json"):     # newest capture first
        tpath = os.path.join(ROOT, "profiles", tname)
        if not os.path.isfile(tpath):
            continue
        tj = json.load(open(tpath))
        if tj.get("workload") == args.workload and tj.get("precision") == args.precision and args.structure == "staircase":
            alias = {"node_transform pair, head fused (tcgen05) + head_finish": "node_transform pair (tcgen05)"}
            for k in kernels:
                if k.get("traffic") is None:
                    k["traffic"] = tj["bytes"].get(k["kernel"], tj["bytes"].get(alias.get(k["kernel"], "")))
    dom = max([k for k in kernels if "unsorted" not in k["kernel"]], key=lambda k: k["ms"])
    roof = {"bound": dom["bound"], "achieved": dom["achieved"], "peak": dom["peak"], "unit": dom["unit"],
            "frac": dom["frac"], "traffic": dom.get("traffic"), "kernel": dom["kernel"], "peak_source": peaks["source"] +
            (" (burst bf16 figure: kernel timed alone, after a 1 s idle and 50 ms between kernels)" if dom["bound"] == "tensor" else " (copy bandwidth)")}
    return {"roofline": roof, "kernels": kernels}


# ----------------------------------------------------------------------------------------------- CPU arms
def _port_setup(cfg, lp):
    from oracle import port        # the one place bench.py executes oracle/: the timed CPU baseline
    torch.manual_seed(0)
    model = port.PortGCN_FC(8, 8, hids=cfg["hids"], depth=cfg["depth"]).eval()
    g = port.graph_from_coo(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n)
    return port, model, g


def cpu_step(port, model, lp):
    """One LP through the reference's CPU path: graph construction (MyToBipartite equivalent), GCN_FC
    forward, inference_gnn -- all on the host cores."""
    g = port.graph_from_coo(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n)
    tg = port.TorchGraph(g)
    with torch.no_grad():
        lc, lv = model(torch.from_numpy(lp.c_feas), torch.from_numpy(lp.v_feas), tg)
        return port.inference_gnn_t(torch.cat((lc, lv), 0), lp.m)


def cpu_baseline(cfg, lp, args, sample_budget_s=20.0):
    import warnings
    warnings.filterwarnings("ignore")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    port, model, _ = _port_setup(cfg, lp)
    cpu_step(port, model, lp)                      # warm-up
    t0 = time.perf_counter()
    reps = 0
    while True:
        cpu_step(port, model, lp)
        reps += 1
        if time.perf_counter() - t0 > sample_budget_s or reps >= 20:
            break
    dt = (time.perf_counter() - t0) / reps
    return {"value": 1.0 / dt, "unit": "LPs/s", "cores": cores, "kind": "port",
            "sample": f"{reps} LPs of the same workload ({cfg['name']}, fp32), {dt * 1e3:.0f} ms/LP; oracle port of the "
                      f"reference CPU path (PyG-equivalent restatement, not the PyG binary); torch threads={cores}",
            "mp_edges_per_sec": mp_edges(lp.nnz, cfg["depth"]) / dt}


def train_sampled_throughput(cfg, lp, dev, precision, steps, world, seeds_per_batch=16_384, fanout=6):
    """C3's sampled mini-batch variant (reference train.py:105-116: NeighborLoader with num_neighbors=[6]*(depth-1) over an
    LP kept whole on the device): every step = neighbour sampling + induced-subgraph build (csrc/sample.cu,
    graph_build.cu) + forward + balanced loss + backward + gradient all-reduce (N > 1) + Adam on the mini-batch."""
    from lpgnn_b200 import arch
    from lpgnn_b200.graph import BipartiteCSR
    from lpgnn_b200.losses import balanced
    from lpgnn_b200.sampling import NeighborSubgraphLoader, ResidentLP
    from lpgnn_b200.train import allreduce_gradients, broadcast_parameters
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=cfg["hids"], depth=cfg["depth"]).to(dev).train().set_precision(precision)
    broadcast_parameters(model, world)
    params = list(model.parameters())
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=5e-4, fused=True)
    g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n, dev, is_sorted=True)
    res = ResidentLP(g, torch.from_numpy(lp.c_feas).to(dev), torch.from_numpy(lp.v_feas).to(dev),
                     torch.from_numpy(lp.y_s).to(dev), torch.from_numpy(lp.y_t).to(dev))
    hops = max(cfg["depth"] - 1, 1)          # train.py:108-110: depth - 1 hops for GCN_FC (the last layer is the FC head)
    loader = NeighborSubgraphLoader(res, [fanout] * hops, seeds_per_batch, shuffle=True, drop_last=True, seed=1)
    sizes, stamps = [], []

    def run(count):
        done, loss = 0, None
        while done < count:
            for batch in loader:
                batch.to(dev, non_blocking=True)
                lc, lv = model(batch)
                lc, lv = lc[:batch.s_bs], lv[:batch.t_bs]
                loss = balanced(lc, lv, batch.y_s[:batch.s_bs], batch.y_t[:batch.t_bs])
                opt.zero_grad(set_to_none=True)
                loss.backward()
                allreduce_gradients(params, world)
                opt.step()
                sizes.append((batch.x_s.shape[0] + batch.x_t.shape[0], batch.edge_index.nnz()))
                stamps.append(time.perf_counter())
                done += 1
                if done >= count:
                    break
        return loss

    run(10)          # mini-batches differ in size: let the caching allocator see the range before the timed region
    barrier(world)
    torch.cuda.synchronize()
    del sizes[:]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    allocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    reserved0 = torch.cuda.memory_reserved(dev)
    del stamps[:]
    stamps.append(time.perf_counter())
    e0.record()
    loss = run(steps)
    e1.record()
    torch.cuda.synchronize()
    device_allocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - allocs0
    barrier(world)
    t_ms = max_over_ranks(e0.elapsed_time(e1), world, dev)
    assert bool(torch.isfinite(loss))
    nodes, nnz = float(np.mean([a for a, _ in sizes])), float(np.mean([b for _, b in sizes]))
    sps = world * steps / (t_ms / 1e3)
    return {"workload": f"{cfg['name']}-shaped LP resident on the device, mini-batches of {loader.batch_size} seed nodes, fan-out "
                        f"[{fanout}]*{hops}: sampling + induced subgraph + training step"
                        f"{' + NCCL grad all-reduce' if world > 1 else ''}",
            "precision": precision, "steps": steps, "ms_per_step": t_ms / steps, "minibatches_per_sec": sps,
            "seed_nodes_per_sec": sps * loader.batch_size, "mean_sampled_nodes": nodes, "mean_sampled_nnz": nnz,
            "cudaMalloc_calls_in_timed_region": int(device_allocs),
            "reserved_bytes_growth_in_timed_region": int(torch.cuda.memory_reserved(dev) - reserved0),
            "host_ms_per_step_median_max": [float(np.median(np.diff(stamps)) * 1e3), float(np.max(np.diff(stamps)) * 1e3)],
            "mp_edges_per_sec_fwd_bwd": mp_edges(nnz, cfg["depth"], fwd_bwd=True) * sps}


def cpu_sweep_baseline(lps, hids, sample_budget_s=15.0):
    """C5: the same sweep order on the oracle port of the reference CPU path, all host threads, until the time
    budget is spent (a bounded sample: the first LPs of the population)."""
    import warnings
    warnings.filterwarnings("ignore")
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    port, model, _ = _port_setup(dict(hids=hids, depth=3), lps[0])
    cpu_step(port, model, lps[0])
    t0, reps = time.perf_counter(), 0
    while reps < len(lps) and time.perf_counter() - t0 < sample_budget_s:
        cpu_step(port, model, lps[reps])
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    nnz_mean = float(np.mean([lp.nnz for lp in lps[:reps]]))
    return {"value": 1.0 / dt, "unit": "LPs/s", "cores": cores, "kind": "port",
            "sample": f"the first {reps} LPs of the sweep (mean nnz {nnz_mean:.0f}), fp32, {dt * 1e3:.0f} ms/LP; oracle port of "
                      f"the reference CPU path (PyG-equivalent restatement, not the PyG binary); torch threads={cores}",
            "mp_edges_per_sec": mp_edges(nnz_mean, 3) / dt}


def cpu_train_baseline(cfg, lp, sample_budget_s=8.0, max_steps=200, warm=3):
    """configs[0] of BASELINE.json: the reference's CPU training step (train.py:117-129: forward, balanced loss,
    backward, Adam) on the oracle port, all host threads; a bounded sample of steps on the same LP."""
    import warnings
    warnings.filterwarnings("ignore")
    from oracle import port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    model = port.PortGCN_FC(8, 8, hids=cfg["hids"], depth=cfg["depth"]).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
    tg = port.TorchGraph(port.graph_from_coo(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n))
    x_s, x_t = torch.from_numpy(lp.c_feas), torch.from_numpy(lp.v_feas)
    y_s, y_t = torch.from_numpy(lp.y_s), torch.from_numpy(lp.y_t)

    def step():
        lc, lv = model(x_s, x_t, tg)
        loss = port.balanced_loss(lc, lv, y_s, y_t)
        opt.zero_grad()
        loss.backward()
        opt.step()

    for _ in range(warm):
        step()
    t0, reps = time.perf_counter(), 0
    while reps < max_steps and (reps == 0 or time.perf_counter() - t0 < sample_budget_s):
        step()
        reps += 1
    dt = (time.perf_counter() - t0) / reps
    return {"value": 1.0 / dt, "unit": "LPs/s", "cores": cores, "kind": "port", "ms_per_step": dt * 1e3,
            "sample": f"{reps} training steps (fwd + balanced loss + bwd + Adam) on one {cfg['name']}-shaped LP, fp32; oracle port "
                      f"of the reference CPU path (PyG-equivalent restatement, not the PyG binary); torch threads={cores}",
            "mp_edges_per_sec_fwd_bwd": mp_edges(lp.nnz, cfg["depth"], fwd_bwd=True) / dt}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle port; the Python
    reference cannot travel to the GPU box), all host threads, same config / metric / unit."""
    import warnings
    warnings.filterwarnings("ignore")
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if args.workload == "C5":
        # the sweep population of the GPU arm (rank 0's share at N=1): one LP per step, cycled in the same order
        cfg = dict(name="C5", hids=args.sweep_hids, depth=3)
        pop = synth.lp_population(args.sweep_distinct, seed=1239)
        lps_list = [synth.processed_lp(m, n, z, seed=sd) for (m, n, z, sd) in pop]
    else:
        cfg = workload_spec(args.workload)
        ring = max(1, args.distinct if cfg["nnz"] <= 2_000_000 else 1)       # rank 0's ring of the GPU arm
        lps_list = [synth.processed_lp(cfg["m"], cfg["n"], cfg["nnz"], seed=cfg["seed"] + 17 * i, structure=args.structure)
                    for i in range(ring)]
    lp = lps_list[0]
    port, model, _ = _port_setup(cfg, lp)
    # the requested warm-up, bounded in time like the timed steps below (C2: ~0.6 s per CPU step)
    warm_done, tw = 0, time.perf_counter()
    while warm_done < max(args.warmup, 1) and (warm_done == 0 or time.perf_counter() - tw < args.cpu_cap_seconds / 3):
        cpu_step(port, model, lps_list[warm_done % len(lps_list)])
        warm_done += 1
    # bounded: a driver-chosen K sized for the GPU arm must not turn into tens of minutes of CPU work
    requested, done = args.steps, 0
    t0 = time.perf_counter()
    while done < requested and (done == 0 or time.perf_counter() - t0 < args.cpu_cap_seconds):
        cur = lps_list[done % len(lps_list)]
        status = cpu_step(port, model, cur)
        assert int((status == 1).sum()) == cur.m
        done += 1
    dt = time.perf_counter() - t0
    args.steps = done
    lps = args.steps / dt
    nnz_mean = float(np.mean([x.nnz for x in lps_list]))
    print(json.dumps({
        "impl": "reference",
        "metric": "LPs/sec (basis prediction: graph build + GCN_FC forward + basis selection)",
        "value": lps, "unit": "LPs/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
        "warmup": warm_done, "warmup_requested": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": (sweep_workload_string(len(lps_list), nnz_mean, cfg["hids"]) if args.workload == "C5" else
                                workload_string(cfg["name"], cfg["hids"], cfg["depth"], args.structure, lp.m, lp.n, lp.nnz)),
                   "precision": "fp32", "structure": args.structure,
                   "note": "rank 0's first LP of the GPU arm's ring (same generator and seed); one host, all cores"},
        "mp_edges_per_sec": mp_edges(nnz_mean, cfg["depth"]) * lps,
        "cpu_baseline": {"value": lps, "unit": "LPs/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} LPs, one per step; oracle port of the reference CPU path "
                                   f"(reference arch.py/val.py restated with torch CPU ops; PyG/torch_sparse are not "
                                   f"installable here), torch threads={cores}"},
        "e2e": {"value": lps, "unit": "LPs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "steps_requested": requested,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C3", "C4", "C5"])
    ap.add_argument("--train-steps", type=int, default=0, help="timed steps of the training sub-arm (default: min(steps, 50))")
    ap.add_argument("--sweep-distinct", type=int, default=192, help="C5: distinct LPs materialised per GPU (cycled)")
    ap.add_argument("--sweep-hids", type=int, default=1024)
    ap.add_argument("--sweep-pack", type=int, default=1, help="C5 e2e arm: pack LPs block-diagonally (1) or one call per LP (0)")
    # fp16 (IEEE half storage, fp32 accumulate) is the 16-bit mode that passes every parity gate of the north star
    # (>= 99.9 % status agreement); bf16 runs the same kernels at the same rate and is reported beside it
    # (default: fp16 for every workload; C4 -- "bf16 full-graph inference" in BASELINE.json -- headlines the gate-passing
    # 16-bit mode too and carries the bf16 figure in the same line)
    ap.add_argument("--precision", default=None, choices=["bf16", "fp16", "fp32", "fp32_simt"])
    ap.add_argument("--distinct", type=int, default=4, help="distinct LPs per GPU cycled by the timed loops (C1-C3)")
    ap.add_argument("--no-precisions", action="store_true", help="skip the fp32 / other 16-bit sub-measurements")
    ap.add_argument("--no-sweep", action="store_true", help="skip the C5 packed-sweep sub-measurement of the default run")
    ap.add_argument("--structure", default="staircase", choices=["staircase", "uniform"])
    ap.add_argument("--inflight", type=int, default=3, help="LPs in flight on alternating streams (both arms)")
    ap.add_argument("--kernel-reps", type=int, default=10)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--cpu-cap-seconds", type=float, default=180.0,
                    help="--impl reference: stop after this much timed CPU work even if fewer than --steps LPs ran")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-kernels", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    if args.precision is None:
        args.precision = "fp16"
    if args.impl == "reference":
        args.steps = args.steps if args.steps is not None else 3
        args.warmup = args.warmup if args.warmup is not None else 1
        run_reference(args)
    elif args.workload == "C5":
        args.steps = args.steps if args.steps is not None else 1000
        args.warmup = args.warmup if args.warmup is not None else 10
        run_sweep(args)
    else:
        args.steps = args.steps if args.steps is not None else 300
        args.warmup = args.warmup if args.warmup is not None else 10
        run_gpu(args)


if __name__ == "__main__":
    main()
