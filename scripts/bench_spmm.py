"""Times every kernel setting of the wide-feature SpMM (lpgnn_spmm_ex: slab_bytes, unroll) on a C2/C4-shaped LP and checks
each is bit-identical to the row-per-warp kernel.  Usage: python scripts/bench_spmm.py [C2|C4] [staircase|uniform] [bf16|fp32]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lpgnn_b200
from lpgnn_b200 import _lib, synth
from lpgnn_b200.graph import BipartiteCSR

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
structure = sys.argv[2] if len(sys.argv) > 2 else "staircase"
dt = torch.bfloat16 if (len(sys.argv) <= 3 or sys.argv[3] == "bf16") else torch.float32
H = 1024
dev = torch.device("cuda:0")
lp = synth.config_lp(cfg, structure)
g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype("float32"), lp.m, lp.n, dev, is_sorted=True)
csr, csc = g.views()
torch.manual_seed(0)
left = torch.randn(lp.m, H, device=dev).to(dt)
right = torch.randn(lp.n, H, device=dev).to(dt)
lib = _lib.load()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
s = 2 if dt == torch.bfloat16 else 4
alg = 2 * (lp.m + lp.n) * H * s + 16 * lp.nnz + 4 * (lp.m + lp.n + 2)


def run(view, x, y, slab, unroll):
    ptr_, idx, val, rows = view
    rc = lib.lpgnn_spmm_ex(ptr_.data_ptr(), idx.data_ptr(), val.data_ptr(), rows, x.data_ptr(), y.data_ptr(), H,
                           _lib.dtype_code(dt), slab, unroll, _lib.stream_ptr())
    _lib.check(rc, "spmm_ex")


def timeit(f, n=15):
    for _ in range(3):
        f()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); f(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


ys, yt = torch.empty_like(left), torch.empty_like(right)
run(csr, right, ys, -1, 0); run(csc, left, yt, -1, 0)
ref_s, ref_t = ys.clone(), yt.clone()
print(f"{cfg} {structure} {dt} m={lp.m} n={lp.n} nnz={lp.nnz} algorithmic bytes/pair={alg/1e6:.0f} MB")
configs = [(-1, 0), (0, 0), (512, 2), (512, 4), (1024, 2), (1024, 4)]
if os.environ.get("SPMM_CONFIGS"):
    configs = [tuple(int(v) for v in c.split(",")) for c in os.environ["SPMM_CONFIGS"].split(";")]
if os.environ.get("SPMM_NCU"):          # one launch per direction and setting, for a profiler capture
    for cfg_ in configs:
        flush.zero_(); run(csr, right, ys, *cfg_)
        flush.zero_(); run(csc, left, yt, *cfg_)
    torch.cuda.synchronize()
    sys.exit(0)
for c_ in configs:
    ys.zero_(); yt.zero_()
    run(csr, right, ys, *c_); run(csc, left, yt, *c_)
    ok = torch.equal(ys, ref_s) and torch.equal(yt, ref_t)
    t_s = timeit(lambda: run(csr, right, ys, *c_))
    t_t = timeit(lambda: run(csc, left, yt, *c_))
    t = t_s + t_t
    print(f"slab={c_[0]:5d} unroll={c_[1]}: A.R {t_s*1e3:7.1f} us  At.L {t_t*1e3:7.1f} us  pair {t*1e3:7.1f} us "
          f"= {alg/t/1e6:7.0f} GB/s ({alg/t/1e6/6538.9:.3f} of HBM)  bit-identical={ok}", flush=True)
