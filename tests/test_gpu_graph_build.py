"""(a1) graph construction on the device must be BIT-EXACT with the oracle's ordering
(oracle.port.graph_from_coo restates reference dataset.py:301-304 + arch.py:71)."""
import numpy as np
import pytest
import torch

from conftest import make_graph_arrays
from oracle import port

pytestmark = pytest.mark.gpu


def _build(row, col, val, m, n, dev, i64=False):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200.graph import BipartiteCSR
    ei = torch.stack([torch.from_numpy(row), torch.from_numpy(col)])
    g = BipartiteCSR.from_edge_index(ei, torch.from_numpy(val), (m, n)).to(dev)
    torch.cuda.synchronize()
    return g


def _assert_equal(g, ref):
    c = lambda t: t.cpu().numpy()
    np.testing.assert_array_equal(c(g.rowptr), ref.rowptr.astype(np.int32))
    np.testing.assert_array_equal(c(g.col), ref.col.astype(np.int32))
    np.testing.assert_array_equal(c(g.val).view(np.uint32), ref.val.view(np.uint32))
    np.testing.assert_array_equal(c(g.colptr), ref.colptr.astype(np.int32))
    np.testing.assert_array_equal(c(g.row_csc), ref.row_csc.astype(np.int32))
    np.testing.assert_array_equal(c(g.val_csc).view(np.uint32), ref.val_csc.view(np.uint32))
    np.testing.assert_array_equal(c(g.csr2csc), ref.csr2csc.astype(np.int32))


@pytest.mark.parametrize("m,n,z,seed", [
    (5, 7, 12, 0), (1, 1, 1, 1), (1000, 2000, 10_000, 2), (300, 70_000, 40_000, 3),
    (70_000, 300, 40_000, 4), (50_000, 100_000, 500_000, 5), (4097, 4097, 4096 * 3 + 1, 6)])
def test_build_matches_oracle(cuda, m, n, z, seed):
    row, col, val = make_graph_arrays(m, n, z, seed)
    g = _build(row, col, val, m, n, cuda)
    _assert_equal(g, port.graph_from_coo(row, col, val, m, n))


def test_build_sorted_input_and_empty_rows(cuda):
    m, n = 2000, 3000
    row, col, val = make_graph_arrays(m, n, 5000, 7, sort=True)
    keep = (row % 3 != 0) & (col % 5 != 0)          # whole rows / columns without entries
    row, col, val = row[keep], col[keep], val[keep]
    g = _build(row, col, val, m, n, cuda)
    _assert_equal(g, port.graph_from_coo(row, col, val, m, n))


def test_build_duplicates_keep_input_order(cuda):
    m, n = 50, 60
    row, col, val = make_graph_arrays(m, n, 4000, 8, dup=True)
    g = _build(row, col, val, m, n, cuda)
    _assert_equal(g, port.graph_from_coo(row, col, val, m, n))


def test_build_empty_graph(cuda):
    m, n = 10, 20
    e = np.zeros(0, dtype=np.int64)
    g = _build(e, e, np.zeros(0, dtype=np.float32), m, n, cuda)
    assert g.nnz() == 0
    assert g.rowptr.cpu().tolist() == [0] * (m + 1)
    assert g.colptr.cpu().tolist() == [0] * (n + 1)


def test_build_dense_row_and_column(cuda):
    m, n = 3000, 5000
    row = np.concatenate([np.full(n, 7), np.arange(m)]).astype(np.int64)      # row 7 dense, column 11 dense
    col = np.concatenate([np.arange(n), np.full(m, 11)]).astype(np.int64)
    key, first = np.unique(row * n + col, return_index=True)
    row, col = row[first], col[first]
    val = np.random.default_rng(9).uniform(-1, 1, row.shape[0]).astype(np.float32)
    perm = np.random.default_rng(10).permutation(row.shape[0])
    row, col, val = row[perm], col[perm], val[perm]
    g = _build(row, col, val, m, n, cuda)
    _assert_equal(g, port.graph_from_coo(row, col, val, m, n))


def test_full_size_properties_c4_shape(cuda):
    """BASELINE C4 shape (1M x 2M, ~1e7 nnz): size-independent properties instead of the oracle."""
    m, n, z = 1_000_000, 2_000_000, 10_000_000
    rng = np.random.default_rng(11)
    row = rng.integers(0, m, size=z)
    col = rng.integers(0, n, size=z)
    val = rng.uniform(-1, 1, size=z).astype(np.float32)
    g = _build(row.astype(np.int64), col.astype(np.int64), val, m, n, cuda)
    rowptr, colptr = g.rowptr.long(), g.colptr.long()
    assert int(rowptr[0]) == 0 and int(rowptr[-1]) == z and int(colptr[-1]) == z
    assert bool((rowptr[1:] >= rowptr[:-1]).all()) and bool((colptr[1:] >= colptr[:-1]).all())
    r, c, v = g.coo()
    key = r * n + c
    assert bool((key[1:] >= key[:-1]).all())                       # CSR canonical order
    rt, ct, vt = g.t().coo()                                       # rows of A^T = columns of A
    keyt = rt * m + ct
    assert bool((keyt[1:] >= keyt[:-1]).all())                     # CSC canonical order
    assert torch.equal(g.val[g.csr2csc.long()], g.val_csc)         # permutation consistency
    assert torch.equal(torch.sort(g.csr2csc)[0], torch.arange(z, device=cuda, dtype=torch.int32))
    # multiset of entries preserved: checksums of (row,col,val-bits) triples
    bits = torch.from_numpy(val.view(np.int32).astype(np.int64)).to(cuda)
    chk_in = (torch.from_numpy(row).to(cuda) * 1_000_003 + torch.from_numpy(col).to(cuda) * 7919 + bits).sum()
    chk_out = (r * 1_000_003 + c * 7919 + g.val.view(torch.int32).long()).sum()
    assert int(chk_in) == int(chk_out)


def test_build_sorted_hint_matches_and_is_verified(cuda):
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200.graph import BipartiteCSR
    m, n = 3000, 7000
    row, col, val = make_graph_arrays(m, n, 30_000, 12, sort=True)
    g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, cuda, is_sorted=True).check()
    _assert_equal(g, port.graph_from_coo(row, col, val, m, n))
    # a false claim is detected on the device and reported by check()
    perm = np.random.default_rng(0).permutation(row.shape[0])
    bad = BipartiteCSR.from_coo_arrays(row[perm], col[perm], val[perm], m, n, cuda, is_sorted=True)
    with pytest.raises(ValueError, match="not in"):
        bad.check()
    # out-of-range indices given on the device are reported too
    r = torch.tensor([0, 1, m], dtype=torch.int32, device=cuda)
    c = torch.tensor([0, 1, 2], dtype=torch.int32, device=cuda)
    with pytest.raises(ValueError, match="out of range"):
        BipartiteCSR.from_coo(r, c, torch.ones(3, device=cuda), m, n).check()


def test_sparse_tensor_api_subset(cuda):
    """The SparseTensor methods the reference calls on batch.edge_index."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200.graph import BipartiteCSR
    m, n = 40, 70
    row, col, val = make_graph_arrays(m, n, 500, 13)
    g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, cuda)
    dense = np.zeros((m, n), dtype=np.float64)
    dense[row, col] = val
    assert g.nnz() == row.shape[0] and g.sparse_sizes() == (m, n) and g.t().sparse_sizes() == (n, m)
    assert abs(g.density() - row.shape[0] / (m * n)) < 1e-12
    np.testing.assert_allclose(g.sum(1).cpu().numpy(), dense.sum(1), atol=1e-5)
    np.testing.assert_allclose(g.sum(0).cpu().numpy(), dense.sum(0), atol=1e-5)
    ones = g.clone().set_value(torch.ones(g.nnz()), layout="coo")       # dataset.py:135-138
    np.testing.assert_array_equal(ones.sum(0).cpu().numpy(), (dense != 0).sum(0))
    np.testing.assert_array_equal(ones.sum(1).cpu().numpy(), (dense != 0).sum(1))
    assert torch.equal(g.storage.value(), g.val) and torch.equal(g.t().storage.value(), g.val_csc)
    r, c, v = g.t().coo()
    np.testing.assert_array_equal(dense.T[r.cpu().numpy(), c.cpu().numpy()].astype(np.float32), v.cpu().numpy())


def test_mean_normalised_build_matches_oracle_and_feeds_the_model(cuda):
    """LPGNN_GRAPH_MEAN (north_star "degree normalisation"; OFF in the reference): both orientations'
    values divided by the destination degree, bit-exact with the oracle; the forward pass on the normalised graph
    matches the oracle model on the same graph; training refuses it."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import arch, synth
    from lpgnn_b200.data import Data
    from lpgnn_b200.graph import BipartiteCSR
    lp = synth.processed_lp(700, 1500, 7500, seed=17)
    val = lp.a_data.astype(np.float32)
    for is_sorted in (True, False):
        row, col, v = (lp.row, lp.col, val) if is_sorted else (lp.row[::-1].copy(), lp.col[::-1].copy(), val[::-1].copy())
        g = BipartiteCSR.from_coo_arrays(row, col, v, lp.m, lp.n, cuda, is_sorted=is_sorted, normalize="mean")
        torch.cuda.synchronize()
        _assert_equal(g.check(), port.graph_from_coo(row, col, v, lp.m, lp.n, normalize="mean"))
    ref_g = port.graph_from_coo(lp.row, lp.col, val, lp.m, lp.n, normalize="mean")
    assert not np.array_equal(ref_g.val, port.graph_from_coo(lp.row, lp.col, val, lp.m, lp.n).val)
    torch.manual_seed(0)
    model = arch.GCN_FC(8, 8, hids=64, depth=3).to(cuda).eval()
    torch.manual_seed(0)
    ref = port.PortGCN_FC(8, 8, hids=64, depth=3).eval()
    xs, xt = torch.from_numpy(lp.c_feas), torch.from_numpy(lp.v_feas)
    with torch.no_grad():
        lc, lv = model(Data(x_s=xs.to(cuda), x_t=xt.to(cuda), edge_index=g))
        ec, ev = ref(xs, xt, port.TorchGraph(ref_g))
    assert float((lc.cpu() - ec).abs().max()) < 1e-3 and float((lv.cpu() - ev).abs().max()) < 1e-3
    model.train()
    with pytest.raises(NotImplementedError):
        model(Data(x_s=xs.to(cuda), x_t=xt.to(cuda), edge_index=g))


@pytest.mark.parametrize("m,n,z,seed", [(5, 7, 12, 0), (1, 1, 1, 1), (3000, 7000, 30_000, 12), (50_000, 100_000, 500_000, 5),
                                         (300, 70_000, 40_000, 3), (70_000, 300, 40_000, 4), (400_000, 900_000, 3_000_000, 8)])
def test_one_launch_sorted_build_equals_the_launch_chain(cuda, m, n, z, seed):
    """Sorted COO -> CSR + CSC in its three forms -- the plain chain (default: 2 + 3 launches per radix pass), the compact
    chain (1 + 2 per pass: histograms and the CSC payload ride in the scatters) and one cooperative launch (grid barriers
    between the radix phases): bit-identical outputs, the same verification of a false sorted claim / out-of-range
    indices; sizes above the co-resident tile capacity stride over the tiles."""
    import lpgnn_b200  # noqa: F401
    from lpgnn_b200 import _lib
    from lpgnn_b200.graph import BipartiteCSR
    lib = _lib.load()
    row, col, val = make_graph_arrays(m, n, z, seed, sort=True)
    outs = []
    for fused, compact, allowed in ((1, 1, (1,)), (0, 0, (5, 8, 11)), (0, 1, (3, 5, 7))):
        prev_f, prev_c = lib.lpgnn_set_graph_fused(fused), lib.lpgnn_set_graph_compact(compact)
        try:
            torch.cuda.synchronize()
            before = lib.lpgnn_launch_count()
            g = BipartiteCSR.from_coo_arrays(row, col, val, m, n, cuda, is_sorted=True).check()
            launches = lib.lpgnn_launch_count() - before
        finally:
            lib.lpgnn_set_graph_fused(prev_f)
            lib.lpgnn_set_graph_compact(prev_c)
        outs.append(g)
        assert launches in allowed, (fused, compact, launches)
    a, b, c = outs
    for name in ("rowptr", "col", "val", "colptr", "row_csc", "val_csc", "csr2csc"):
        assert torch.equal(getattr(c, name), getattr(b, name)), name
    for name in ("rowptr", "col", "val", "colptr", "row_csc", "val_csc", "csr2csc"):
        assert torch.equal(getattr(a, name), getattr(b, name)), name
    _assert_equal(a, port.graph_from_coo(row, col, val, m, n))
    perm = np.random.default_rng(0).permutation(row.shape[0])
    if row.shape[0] > 3:
        bad = BipartiteCSR.from_coo_arrays(row[perm], col[perm], val[perm], m, n, cuda, is_sorted=True)
        with pytest.raises(ValueError, match="not in"):
            bad.check()
