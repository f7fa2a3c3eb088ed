"""(a2) SpMM parity: CUDA kernel vs the oracle's sequential CSR-order accumulation."""
import numpy as np
import pytest
import torch

from conftest import make_graph_arrays
from oracle import port

pytestmark = pytest.mark.gpu


def _graph(m, n, z, seed, dev):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200.graph import BipartiteCSR
    row, col, val = make_graph_arrays(m, n, z, seed)
    g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, dev)
    return g, port.graph_from_coo(row, col, val, m, n)


@pytest.mark.parametrize("F", [4, 8, 16, 64, 100, 128, 1024])
@pytest.mark.parametrize("m,n,z", [(37, 53, 400), (1000, 2000, 10_000)])
def test_spmm_fp32_both_orientations(cuda, m, n, z, F):
    from lpgnn_b200 import ops
    g, ref = _graph(m, n, z, 100 + F, cuda)
    rng = np.random.default_rng(F)
    xr = rng.standard_normal((n, F)).astype(np.float32)
    xl = rng.standard_normal((m, F)).astype(np.float32)
    csr, csc = g.views()
    ys = ops.spmm(csr, torch.from_numpy(xr).to(cuda)).cpu().numpy()
    yt = ops.spmm(csc, torch.from_numpy(xl).to(cuda)).cpu().numpy()
    es = port.spmm_sequential(ref.rowptr, ref.col, ref.val, xr)
    et = port.spmm_sequential(ref.colptr, ref.row_csc, ref.val_csc, xl)
    # same accumulation order; only fma contraction may differ from the CPU -> a few ulp
    np.testing.assert_allclose(ys, es, rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(yt, et, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("F", [8, 64, 128, 1024])
@pytest.mark.parametrize("dt16", [torch.bfloat16, torch.float16])
def test_spmm_bf16(cuda, F, dt16):
    from lpgnn_b200 import ops
    m, n, z = 500, 900, 6000
    g, ref = _graph(m, n, z, 7, cuda)
    x = torch.randn(n, F, generator=torch.Generator().manual_seed(F)).to(dt16)
    csr, _ = g.views()
    y = ops.spmm(csr, x.to(cuda)).float().cpu().numpy()
    e = port.spmm_sequential(ref.rowptr, ref.col, ref.val, x.float().numpy())   # fp32 accumulate of bf16 inputs
    np.testing.assert_allclose(y, e, rtol=1e-2, atol=1e-2)                       # output rounding to bf16
    # and the rounding is the ONLY difference: re-round the oracle
    e_bf = torch.from_numpy(e).to(dt16).float().numpy()
    assert np.mean(y == e_bf) > 0.99


def test_spmm_empty_rows_and_dense_row(cuda):
    from lpgnn_b200 import ops
    from lpgnn_b200.graph import BipartiteCSR
    m, n, F = 64, 3000, 128
    row = np.concatenate([np.full(n, 5), np.array([9, 9, 63])]).astype(np.int64)   # rows 0-4 etc. empty, row 5 dense
    col = np.concatenate([np.arange(n), np.array([1, 2, 2999])]).astype(np.int64)
    val = np.random.default_rng(3).uniform(-1, 1, row.shape[0]).astype(np.float32)
    g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, cuda)
    ref = port.graph_from_coo(row, col, val, m, n)
    x = np.random.default_rng(4).standard_normal((n, F)).astype(np.float32)
    y = ops.spmm(g.views()[0], torch.from_numpy(x).to(cuda)).cpu().numpy()
    e = port.spmm_sequential(ref.rowptr, ref.col, ref.val, x)
    np.testing.assert_allclose(y, e, rtol=1e-5, atol=1e-4)
    assert not y[0].any() and not y[10].any()


def test_spmm_linearity_full_size_c2(cuda):
    """C2 shape, H=1024: linearity A(x+y) = Ax + Ay and agreement with a float64 torch model on a sample of rows."""
    from lpgnn_b200 import ops
    m, n, z, F = 50_000, 100_000, 500_000, 1024
    g, ref = _graph(m, n, z, 21, cuda)
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(n, F, device=cuda, generator=gen)
    y = torch.randn(n, F, device=cuda, generator=gen)
    csr, _ = g.views()
    a, b, c = ops.spmm(csr, x), ops.spmm(csr, y), ops.spmm(csr, x + y)
    assert float((a + b - c).abs().max()) < 1e-4
    rows = np.random.default_rng(1).integers(0, m, 64)
    xc = x.double()
    for r in rows:
        lo, hi = int(ref.rowptr[r]), int(ref.rowptr[r + 1])
        cols = torch.from_numpy(ref.col[lo:hi]).to(cuda)
        w = torch.from_numpy(ref.val[lo:hi]).to(cuda).double()
        e = (w[:, None] * xc[cols]).sum(0)
        assert float((a[r].double() - e).abs().max()) < 1e-4


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("structure", ["staircase", "uniform"])
def test_banded_sweep_kernel_is_bit_identical_to_row_kernel(cuda, dtype, structure):
    """Every kernel setting of lpgnn_spmm_ex (row-per-warp, 512 B / 1 KB slabs, 2 / 4 gathers in flight, automatic)
    returns the same bits, on a banded LP and on one without locality, both orientations; the row kernel itself is
    pinned to the oracle (sequential CSR-order fp32 accumulation) on a sample of rows."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops, synth
    from lpgnn_b200.graph import BipartiteCSR
    m, n, F = 12_000, 24_000, 1024 if dtype != torch.float32 else 512
    c, b_l, A, b_u, l, u = synth.raw_lp(m, n, 5 * n, 31, structure)
    A = A.tocsr()
    A.sort_indices()
    row = np.repeat(np.arange(m), np.diff(A.indptr))
    val = (A.data / 10.0).astype(np.float32)
    g = BipartiteCSR.from_coo_arrays(row, A.indices, val, m, n, cuda, is_sorted=True)
    gen = torch.Generator(device="cuda").manual_seed(5)
    xs = {0: torch.randn(n, F, device=cuda, generator=gen).to(dtype), 1: torch.randn(m, F, device=cuda, generator=gen).to(dtype)}
    for k, view in enumerate(g.views()):
        ref = ops.spmm(view, xs[k], slab_bytes=-1)
        for slab, unroll in ((0, 0), (512, 2), (512, 4), (1024, 2), (1024, 4)):
            got = ops.spmm(view, xs[k], slab_bytes=slab, unroll=unroll)
            assert torch.equal(got.view(torch.int16 if dtype != torch.float32 else torch.int32),
                               ref.view(torch.int16 if dtype != torch.float32 else torch.int32)), (k, slab, unroll)
        ptr_, idx, v, rows = (t.cpu().numpy() if torch.is_tensor(t) else t for t in view)
        pick = np.unique(np.concatenate([[0, rows - 1], np.random.default_rng(k).integers(0, rows, 200)]))
        sub_ptr = np.concatenate([[0], np.cumsum(ptr_[pick + 1] - ptr_[pick])])
        sub_idx = np.concatenate([idx[ptr_[r]:ptr_[r + 1]] for r in pick])
        sub_val = np.concatenate([v[ptr_[r]:ptr_[r + 1]] for r in pick])
        e = port.spmm_sequential(sub_ptr, sub_idx, sub_val, xs[k].float().cpu().numpy())
        got = ref[torch.from_numpy(pick).to(cuda)].float().cpu().numpy()
        if dtype != torch.float32:
            assert np.mean(got == torch.from_numpy(e).to(dtype).float().numpy()) > 0.99
        else:
            np.testing.assert_allclose(got, e, rtol=1e-5, atol=1e-5)


def test_sweep_kernel_ragged_rows(cuda):
    """Rows longer than one 32-entry fetch, empty rows and a row count that is not a multiple of the CTA's 32 warps."""
    from lpgnn_b200 import ops
    from lpgnn_b200.graph import BipartiteCSR
    m, n, F = 10_007, 5_003, 256
    rng = np.random.default_rng(9)
    lens = rng.choice([0, 1, 3, 31, 32, 33, 70, 200], size=m, p=[.2, .3, .3, .05, .05, .05, .04, .01])
    row = np.repeat(np.arange(m), lens)
    col = np.concatenate([np.sort(rng.choice(n, k, replace=False)) for k in lens]).astype(np.int64)
    val = rng.uniform(-1, 1, row.shape[0]).astype(np.float32)
    g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, cuda, is_sorted=True)
    ref = port.graph_from_coo(row, col, val, m, n)
    x = rng.standard_normal((n, F)).astype(np.float32)
    e = port.spmm_sequential(ref.rowptr, ref.col, ref.val, x)
    xd = torch.from_numpy(x).to(cuda)
    for slab, unroll in ((-1, 0), (512, 2), (512, 4), (1024, 2), (1024, 4)):
        y = ops.spmm(g.views()[0], xd, slab_bytes=slab, unroll=unroll).cpu().numpy()
        np.testing.assert_allclose(y, e, rtol=1e-5, atol=1e-4)
        assert not y[lens == 0].any()


def test_sweep_kernel_writes_only_its_output(cuda):
    """Canary regions around Y (compute-sanitizer is not available on the GPU pool): the banded sweep with a row count
    that is not a multiple of its 32 warps, nor of the per-CTA row range, must not touch a byte outside [rows, F]."""
    from lpgnn_b200 import _lib, synth
    from lpgnn_b200.graph import BipartiteCSR
    lib = _lib.load()
    m, n, F = 10_013, 20_011, 512
    c, b_l, A, b_u, l, u = synth.raw_lp(m, n, 5 * n, 41)
    A = A.tocsr(); A.sort_indices()
    row = np.repeat(np.arange(m), np.diff(A.indptr))
    g = BipartiteCSR.from_coo_arrays(row, A.indices, (A.data / 10).astype(np.float32), m, n, cuda, is_sorted=True)
    for dtype in (torch.float32, torch.bfloat16):
        for (ptr_, idx, val, rows), src_rows in zip(g.views(), (n, m)):
            x = torch.randn(src_rows, F, device=cuda).to(dtype)
            pad = 4096
            buf = torch.full((rows * F + 2 * pad,), 7.0, device=cuda).to(dtype)
            y = buf[pad:pad + rows * F]
            for slab in (512, 1024):
                rc = lib.lpgnn_spmm_ex(ptr_.data_ptr(), idx.data_ptr(), val.data_ptr(), rows, x.data_ptr(), y.data_ptr(), F,
                                       _lib.dtype_code(dtype), slab, 0, _lib.stream_ptr())
                _lib.check(rc, "spmm_ex")
                torch.cuda.synchronize()
                assert bool((buf[:pad] == 7.0).all()) and bool((buf[pad + rows * F:] == 7.0).all())


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_c4_shape_banded_sweep_is_checked_not_only_timed(cuda, dtype):
    """BASELINE config C4's own generator (1M x 2M staircase LP, ~9.8M nonzeros): the banded sweep that the step runs at
    this size against the row kernel (bit-identical) and against the oracle's sequential CSR-order sum on sampled rows,
    both orientations -- element offsets pass 2^31 (2M rows x 1024 features) on the REAL band structure."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import ops, synth
    from lpgnn_b200.graph import BipartiteCSR
    cfg = synth.CONFIGS["C4"]
    lp = synth.processed_lp(cfg["m"], cfg["n"], cfg["nnz"], seed=cfg["seed"])
    F = 1024 if dtype != torch.float32 else 512
    g = BipartiteCSR.from_coo_arrays(lp.row, lp.col, lp.a_data.astype(np.float32), lp.m, lp.n, cuda, is_sorted=True).check()
    gen = torch.Generator(device="cuda").manual_seed(9)
    for k, view in enumerate(g.views()):
        src_rows = lp.n if k == 0 else lp.m
        x = torch.randn(src_rows, F, device=cuda, generator=gen).to(dtype)
        y = ops.spmm(view, x)                              # automatic: the banded sweep at this size
        ref = ops.spmm(view, x, slab_bytes=-1)             # row-per-warp kernel
        assert torch.equal(y.view(torch.int32 if dtype == torch.float32 else torch.int16),
                           ref.view(torch.int32 if dtype == torch.float32 else torch.int16)), k
        del ref
        ptr_, idx, v, rows = view
        pick = torch.from_numpy(np.unique(np.concatenate([[0, rows - 1], np.random.default_rng(k).integers(0, rows, 64)]))).to(cuda)
        beg, end = ptr_[pick].cpu().numpy(), ptr_[pick + 1].cpu().numpy()
        sub_ptr = np.concatenate([[0], np.cumsum(end - beg)])
        sub_idx = np.concatenate([idx[b:e].cpu().numpy() for b, e in zip(beg, end)])
        sub_val = np.concatenate([v[b:e].cpu().numpy() for b, e in zip(beg, end)])
        uniq, inv = np.unique(sub_idx, return_inverse=True)                      # only the gathered source rows travel to the host
        xs = x[torch.from_numpy(uniq).to(cuda)].float().cpu().numpy()
        e = port.spmm_sequential(sub_ptr, inv.astype(sub_idx.dtype), sub_val, xs)
        got = y[pick].float().cpu().numpy()
        if dtype == torch.float32:
            np.testing.assert_allclose(got, e, rtol=1e-5, atol=1e-5)
        else:
            assert np.mean(got == torch.from_numpy(e).to(dtype).float().numpy()) > 0.99
            np.testing.assert_allclose(got, e, rtol=2e-3, atol=2e-3)
        del x, y


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("m,n,z,F", [(50_000, 100_000, 500_000, 1024), (30_000, 9_000, 200_000, 256), (9_000, 30_000, 100_000, 512),
                                     (700, 1300, 6000, 64), (5, 40_000, 60_000, 1024), (20_000, 41_000, 0, 1024)])
def test_pair_launch_equals_two_aggregations(cuda, dtype, m, n, z, F):
    """lpgnn_spmm_pair (A R and A^T L in one launch where both take the banded sweep, the SMs split by work; two plain
    launches otherwise) returns the bits of two lpgnn_spmm calls, with and without the nonzero count as a hint."""
    from lpgnn_b200 import ops
    g, _ = _graph(m, n, z, 3, cuda)
    csr, csc = g.views()
    gen = torch.Generator(device="cuda").manual_seed(m + F)
    left = torch.randn(m, F, device=cuda, generator=gen).to(dtype)
    right = torch.randn(n, F, device=cuda, generator=gen).to(dtype)
    agg_s, agg_t = ops.spmm(csr, right), ops.spmm(csc, left)
    for hint in (g.nnz(), -1):
        ps, pt = ops.spmm_pair(csr, csc, left, right, nnz=hint)
        assert torch.equal(ps, agg_s) and torch.equal(pt, agg_t)


@pytest.mark.parametrize("m,n,z,F", [(50_000, 100_000, 500_000, 1024), (700, 1300, 6000, 128), (30_000, 70_000, 300_000, 256)])
def test_x2_pair_launch_equals_two_aggregations(cuda, m, n, z, F):
    """lpgnn_spmm_x2_pair == two lpgnn_spmm_x2 calls (hi / lo operands and row scales), sweep and two-step shapes."""
    from lpgnn_b200 import ops
    g, _ = _graph(m, n, z, 4, cuda)
    csr, csc = g.views()
    gen = torch.Generator(device="cuda").manual_seed(n + F)
    left = torch.randn(m, F, device=cuda, generator=gen)
    right = torch.randn(n, F, device=cuda, generator=gen)
    # row-scale bounds as the producer of the features reports them: |X[j,:]| <= scale[j] * 2^12
    sl = (left.abs().amax(1) / 4096).clamp_min(1e-30)
    sr = (right.abs().amax(1) / 4096).clamp_min(1e-30)
    (s_ops, s_sc), (t_ops, t_sc) = ops.spmm_x2(csr, right, sr), ops.spmm_x2(csc, left, sl)
    ((hs, ls), ss), ((ht, lt), st) = ops.spmm_x2_pair(csr, csc, left, right, sl, sr, nnz=g.nnz())
    assert torch.equal(hs, s_ops[0]) and torch.equal(ls, s_ops[1]) and torch.equal(ss, s_sc)
    assert torch.equal(ht, t_ops[0]) and torch.equal(lt, t_ops[1]) and torch.equal(st, t_sc)
